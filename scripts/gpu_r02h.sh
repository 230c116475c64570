#!/bin/bash
# round 2, call H: patch schedule knobs - ring depths, L2 prefetch distance, resident weights without chunked tile ranges
O=gpurun_out/r02h; mkdir -p $O
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --verbose --top 60 --no-cpu-baseline --no-eager-baseline --no-extra-configs > $O/bench_$name.json 2> $O/bench_$name.err
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
run base A=1
run a6b4 WFD_PATCH_ASTAGES=6 WFD_PATCH_BSTAGES=4
run a5b6 WFD_PATCH_ASTAGES=5 WFD_PATCH_BSTAGES=6
run a3b8 WFD_PATCH_ASTAGES=3 WFD_PATCH_BSTAGES=8
run pf2 WFD_PATCH_PREFETCH=2
run pf4 WFD_PATCH_PREFETCH=4
run pf8 WFD_PATCH_PREFETCH=8
run a6b4pf4 WFD_PATCH_ASTAGES=6 WFD_PATCH_BSTAGES=4 WFD_PATCH_PREFETCH=4
run res_strided WFD_PATCH_RESIDENT=1 WFD_PATCH_CHUNKED=0
run res_pf4 WFD_PATCH_RESIDENT=1 WFD_PATCH_PREFETCH=4
