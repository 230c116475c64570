#!/bin/bash
# round 2, call AA: specialised epilogue of the backward convs with the fused GroupNorm-backward reduction (epi_fast 3):
# parity, then A/B against the generic loop and with the fused reduction extended to 1536 / 3072 channels
O=gpurun_out/r02aa; mkdir -p $O
timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_tapgemm_gpu.py -m gpu -q --timeout 100 -x -s -k "small or ice_config or tapgemm" > $O/pytest_quick.log 2>&1; echo "pytest rc=$?"
grep -i "rel-L2\|passed\|failed\|error" $O/pytest_quick.log | tail -8
run() { # name, env...
  name=$1; shift
  env "$@" timeout 120 python bench.py --steps 10 --warmup 3 --verbose --top 80 --no-cpu-baseline --no-eager-baseline --no-extra-configs > $O/bench_$name.json 2> $O/bench_$name.err
  python - $O/bench_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d['roofline']
    print("%-14s ms/step %.3f frac %.3f loss %s"%(sys.argv[2], d['ms_per_step'], r['frac'], d['loss_check'][:1]), {k:round(v['ms_per_step'],3) for k,v in r['classes'].items()})
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
run fast3 A=1
run generic WFD_NO_EPI_FAST=3
run fast3_1536 WFD_FUSE_BWD_MAX_C=1536
run fast3_3072 WFD_FUSE_BWD_MAX_C=3072
run fast3b A=1
grep "gn_bwd" $O/bench_fast3.err $O/bench_fast3_1536.err $O/bench_fast3_3072.err
