#!/bin/bash
# round 2, call AG: every epilogue form as its own kernel instantiation (no spills left): whole suite, class times
O=gpurun_out/r02ag; mkdir -p $O
timeout 700 python -m pytest tests -m gpu -x -q --timeout 120 > $O/pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $O/pytest.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 120 python bench.py --steps 10 --warmup 3 --verbose --top 80 --no-cpu-baseline --no-eager-baseline --no-extra-configs > $O/bench_$name.json 2> $O/bench_$name.err
  python - $O/bench_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d['roofline']
    print("%-10s ms/step %.3f frac %.3f mhz %s loss %s"%(sys.argv[2], d['ms_per_step'], r['frac'], d['clocks']['sm_mhz'], d['loss_check'][:1]), {k:round(v['ms_per_step'],3) for k,v in r['classes'].items()})
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
run split_all A=1
run no_fast WFD_NO_EPI_FAST=1
run split_all2 A=1
for f in 0 1; do
timeout 200 python bench.py --batch 1 --latent 256 --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-extra-configs 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('256x256 one GPU ms/step %.3f loss %s'%(d['ms_per_step'], d['loss_check'][:1]))"
done
