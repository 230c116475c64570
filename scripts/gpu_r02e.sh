#!/bin/bash
# round 2, call E: the patch schedule - unit tests first (they separate descriptor mistakes from the rest), then parity + bench
O=gpurun_out/r02e; mkdir -p $O
timeout 600 python -m pytest tests/test_tapgemm_gpu.py -m gpu -q --timeout 300 -k "conv3x3" > $O/pytest_unit.log 2>&1; echo "unit rc=$?"; tail -15 $O/pytest_unit.log
timeout 900 python -m pytest tests/test_tapgemm_gpu.py tests/test_parity_gpu.py -m gpu -x -q --timeout 600 > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --verbose --top 60 --no-cpu-baseline --no-extra-configs > $O/bench_$name.json 2> $O/bench_$name.err
  python - $O/bench_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d['roofline']
    print("%-14s ms/step %.3f frac %.3f dense-eq %.3f loss %s"%(sys.argv[2], d['ms_per_step'], r['frac'], r['frac_dense_equivalent'], d['loss_check']))
    print("   ", {k:(round(v['ms_per_step'],3), round(v['frac'] or 0,3)) for k,v in r['classes'].items()})
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
run patch A=1
run nopatch WFD_NO_PATCH=1
run patch_nores WFD_PATCH_NO_RESIDENT=1
tail -50 $O/bench_patch.err
