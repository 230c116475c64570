#!/bin/bash
# round 2, call D: the lean MMA-issue loop - unit + parity tests, then bench with the grouped-conv knobs
O=gpurun_out/r02d; mkdir -p $O
timeout 900 python -m pytest tests/test_tapgemm_gpu.py tests/test_parity_gpu.py -m gpu -x -q --timeout 600 > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.log
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --steps 10 --warmup 3 --verbose --top 60 --no-cpu-baseline --no-extra-configs > $O/bench_$name.json 2> $O/bench_$name.err
  python - $O/bench_$name.json $name <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=d['roofline']
print("%-14s ms/step %.3f frac %.3f dense-eq %.3f loss %s"%(sys.argv[2], d['ms_per_step'], r['frac'], r['frac_dense_equivalent'], d['loss_check']))
print("   ", {k:(round(v['ms_per_step'],3), round(v['frac'] or 0,3)) for k,v in r['classes'].items()})
PY
}
run base A=1
run hbox8 WFD_REUSE_HBOX=8
run hbox4 WFD_REUSE_HBOX=4
run no2cta WFD_REUSE_2CTA=0
run no2cta_hbox8 WFD_REUSE_2CTA=0 WFD_REUSE_HBOX=8
tail -62 $O/bench_base.err
