#!/bin/bash
# round 2, call AJ: residual pieces of the next tile requested one tile ahead in the plain conv drains: parity, A/B
O=gpurun_out/r02aj; mkdir -p $O
timeout 300 python -m pytest tests/test_tapgemm_gpu.py tests/test_parity_gpu.py -m gpu -q --timeout 100 -x -k "tapgemm or small or ice_config or batch_properties" > $O/pytest_quick.log 2>&1; echo "pytest rc=$?"
tail -2 $O/pytest_quick.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 120 python bench.py --steps 10 --warmup 3 --verbose --top 80 --no-cpu-baseline --no-eager-baseline --no-extra-configs > $O/bench_$name.json 2> $O/bench_$name.err
  python - $O/bench_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d['roofline']
    print("%-10s ms/step %.3f frac %.3f mhz %s loss %s"%(sys.argv[2], d['ms_per_step'], r['frac'], d['clocks']['sm_mhz'], d['loss_check'][:2]), {k:round(v['ms_per_step'],3) for k,v in r['classes'].items()})
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
run prefetch A=1
run upfront WFD_NO_RES_PREFETCH=1
run prefetch2 A=1
grep "side=1" $O/bench_prefetch.err | head -8 | cut -c1-150
grep "side=1" $O/bench_upfront.err | head -8 | cut -c1-150
