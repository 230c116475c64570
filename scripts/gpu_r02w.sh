#!/bin/bash
# round 2, call W: A/B of the fused GroupNorm-backward reduction threshold, tensor-core conv_in^T, epilogue knobs
O=gpurun_out/r02w; mkdir -p $O
run() { # name, env...
  name=$1; shift
  env "$@" timeout 90 python bench.py --steps 10 --warmup 3 --verbose --top 80 --no-cpu-baseline --no-eager-baseline --no-extra-configs > $O/bench_$name.json 2> $O/bench_$name.err
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
run base A=1
run fuse1536 WFD_FUSE_BWD_MAX_C=1536
run fuse3072 WFD_FUSE_BWD_MAX_C=3072
run notclat WFD_NO_TC_LATENT=1
grep "latent\|N=48,BN=48,taps=1" $O/bench_base.err $O/bench_notclat.err
