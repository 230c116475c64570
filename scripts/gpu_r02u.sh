#!/bin/bash
# round 2, call U: 4 / 8 TMEM accumulator stages for narrow N tiles - unit + parity tests, then A/B against two stages
O=gpurun_out/r02u; mkdir -p $O
timeout 300 python -m pytest tests/test_tapgemm_gpu.py tests/test_parity_gpu.py -m gpu -q --timeout 100 -x -k "not config3 and not all_tilesets" > $O/pytest_quick.log 2>&1; echo "pytest rc=$?"
tail -3 $O/pytest_quick.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 90 python bench.py --steps 10 --warmup 3 --verbose --top 60 --no-cpu-baseline --no-eager-baseline --no-extra-configs > $O/bench_$name.json 2> $O/bench_$name.err
  python - $O/bench_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d['roofline']
    print("%-14s ms/step %.3f frac %.3f loss %s"%(sys.argv[2], d['ms_per_step'], r['frac'], d['loss_check'][:1]), {k:round(v['ms_per_step'],3) for k,v in r['classes'].items()}, {k:round(v['ms_per_step'],3) for k,v in r['hbm_kernels'].items()})
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
run acc8 A=1
run acc4 WFD_ACC_STAGES=4
run acc2 WFD_ACC_STAGES=2
