#!/bin/bash
# round 2, call X: softmax backward in the dP epilogue - parity, A/B at 64x64 batch 8 and at 256x256
O=gpurun_out/r02x; mkdir -p $O
timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_tapgemm_gpu.py -m gpu -q --timeout 100 -x -s -k "small or ice_config or two_kernel or tapgemm" > $O/pytest_quick.log 2>&1; echo "pytest rc=$?"
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
run fused A=1
run twokernel WFD_NO_FUSE_ATTN_BWD=1 WFD_NO_FUSE_ATTN_FWD=1
run fwdonly WFD_NO_FUSE_ATTN_BWD=1
run bwdonly WFD_NO_FUSE_ATTN_FWD=1
grep "softmax\|rowdot\|tapgemm\[attn" $O/bench_fused.err
for f in 0 1; do
WFD_NO_FUSE_ATTN_BWD=$f WFD_NO_FUSE_ATTN_FWD=$f timeout 200 python bench.py --batch 1 --latent 256 --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-extra-configs 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('256x256 NO_FUSE_ATTN_BWD=$f ms/step %.3f loss %s'%(d['ms_per_step'], d['loss_check'][:1]))"
done
