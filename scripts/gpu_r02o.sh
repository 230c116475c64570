#!/bin/bash
# round 2, call O: tile order of the grouped convs (N fastest = whole pixel rows read at a time) and the transform's cost split
O=gpurun_out/r02o; mkdir -p $O
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
run unfused_ng64 WFD_NO_FUSE_APPLY=1 WFD_CONV_NGROUP=64
run unfused_ng8 WFD_NO_FUSE_APPLY=1 WFD_CONV_NGROUP=8
run unfused_ng2 WFD_NO_FUSE_APPLY=1 WFD_CONV_NGROUP=2
run fused_ng64 WFD_CONV_NGROUP=64
run fused_noxf WFD_TG_SKIP=16
run fused_a4 WFD_PATCH_ASTAGES=4
run fused_a5 WFD_PATCH_ASTAGES=5
