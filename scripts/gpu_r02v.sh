#!/bin/bash
# round 2, call V: ms per map as a function of the per-launch batch (is a batch of 8 better served as 2 x 4 whose tensors fit L2?)
O=gpurun_out/r02v; mkdir -p $O
for b in 2 4 6 8 12 16; do
  timeout 120 python bench.py --batch $b --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-extra-configs > $O/bench_b$b.json 2> $O/bench_b$b.err
  python - $O/bench_b$b.json $b <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d['roofline']; b=int(sys.argv[2])
    print("batch %2d ms/step %.3f ms/map %.3f maps/s %.1f"%(b, d['ms_per_step'], d['ms_per_step']/b, d['value']), {k:round(v['ms_per_step']/b,4) for k,v in r['hbm_kernels'].items()}, "tapgemm/map %.3f"%(r['kernel_ms_per_step']/b))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
done
timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_pipeline_gpu.py -m gpu -q --timeout 100 -x -k "not all_tilesets" > $O/pytest_quick.log 2>&1; echo "pytest rc=$?"
tail -3 $O/pytest_quick.log
WFD_NO_TC_LATENT=1 timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-extra-configs 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('no_tc_latent ms/step %.3f'%d['ms_per_step'])"
