#!/bin/bash
# quick regression + bench: small-config parity, config-1 golden, then the bench line with the per-kernel profile
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout 300 -x -k "small or ice_config1 or closed_form" 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --verbose --top 40 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print("ms/step %.3f value %.1f e2e %.1f roofline %.3f launches %d"%(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac'],d['gpu_launches']))
PY
tail -44 gpurun_out/bench.err
