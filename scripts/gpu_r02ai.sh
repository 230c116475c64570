#!/bin/bash
# round 2, call AI: fused GroupNorm-backward reduction up to 1536 / 3072 channels again, now that its drain has its own
# spill-free kernel
O=gpurun_out/r02ai; mkdir -p $O
run() { # name, env...
  name=$1; shift
  env "$@" timeout 120 python bench.py --steps 10 --warmup 3 --verbose --top 80 --no-cpu-baseline --no-eager-baseline --no-extra-configs > $O/bench_$name.json 2> $O/bench_$name.err
  python - $O/bench_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d['roofline']
    h=r.get('hbm_kernels',{})
    print("%-10s ms/step %.3f frac %.3f mhz %s loss %s"%(sys.argv[2], d['ms_per_step'], r['frac'], d['clocks']['sm_mhz'], d['loss_check'][:1]), {k:round(v['ms_per_step'],3) for k,v in r['classes'].items()}, 'gn_bwd_reduce %.3f'%h.get('gn_bwd_reduce',{}).get('ms_per_step',0))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
run c768 A=1
run c1536 WFD_FUSE_BWD_MAX_C=1536
run c3072 WFD_FUSE_BWD_MAX_C=3072
run c768b A=1
