#!/bin/bash
for b in 1 8; do for g in "" "--no-graph"; do
python bench.py --batch $b --steps 20 --warmup 3 --no-cpu-baseline $g 2>gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B=$b graph=%s ms/step %.3f value %.1f e2e %.1f (e2e ms %.3f) launches %d'%(d['cuda_graph'],d['ms_per_step'],d['value'],d['e2e']['value'],d['e2e']['ms_per_step'],d['gpu_launches']))" || tail -5 gpurun_out/bench.err
done; done
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_pipeline_gpu.py -m gpu -q --timeout 600 -k "not large_map" 2>&1 | tail -3
