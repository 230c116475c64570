#!/bin/bash
# round 2, call N: GroupNorm+SiLU applied in the consuming conv's operand path (transform warps) - parity, then A/B bench
O=gpurun_out/r02n; mkdir -p $O


timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout 300 -x -s -k "small or ice_config1 or closed_form or standalone" > $O/pytest_quick.log 2>&1; echo "pytest rc=$?"
grep -i "fused vs\|passed\|failed\|error" $O/pytest_quick.log | tail -8
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --verbose --top 60 --no-cpu-baseline --no-eager-baseline --no-extra-configs > $O/bench_$name.json 2> $O/bench_$name.err
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
run fused A=1
run unfused WFD_NO_FUSE_APPLY=1
