#!/bin/bash
# round 2: the driver's 8-GPU launch of bench.py (weak scaling by map + config 3 balanced over the ranks + config 4 in row bands)
O=gpurun_out/r02scale; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 > $O/scale_8.json 2> $O/scale_8.err
echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02scale/scale_8.json').read().strip().splitlines()[-1])
print("n_gpus", d['n_gpus'], "ms/step", d['ms_per_step'], "value", d['value'], "e2e", d['e2e']['value'])
print(json.dumps(d.get('other_configs'))[:1500])
PY
tail -3 $O/scale_8.err
