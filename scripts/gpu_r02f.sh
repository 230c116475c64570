#!/bin/bash
# round 2, call F: patch schedule with the fixed resident-weights handshake; role counters
O=gpurun_out/r02f; mkdir -p $O
timeout 900 python -m pytest tests/test_tapgemm_gpu.py tests/test_parity_gpu.py -m gpu -x -q --timeout 600 > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --verbose --top 60 --no-cpu-baseline --no-eager-baseline --no-extra-configs > $O/bench_$name.json 2> $O/bench_$name.err
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
run patch_nores WFD_PATCH_NO_RESIDENT=1
run patch_noEPI WFD_TG_SKIP=8
run patch_onlyMMA WFD_TG_SKIP=11
run patch_onlyTMA WFD_TG_SKIP=12
grep "reuse=2" $O/bench_patch.err | head -30
WFD_TG_DBG=1 timeout 300 python bench.py --no-graph --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-extra-configs > /dev/null 2> $O/roles.txt
python scripts/tgdbg_report.py $O/roles.txt | head -40
