#!/bin/bash
# weak scaling by map (BASELINE config 2 per GPU) as the driver launches it. Usage: gpu_scale.sh N
cd "$(dirname "$0")/.."
N=${1:-8}
mkdir -p gpurun_out
timeout -k 5 ${SCALE_TIMEOUT:-200} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
echo "rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/scale_$N.json").read().strip().splitlines()[-1])
    print("N=%d ms/step %.3f value %.1f e2e %.1f frac %.3f" % (d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"]))
except Exception as e:
    print("no json", e)
PY
