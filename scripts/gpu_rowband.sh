#!/bin/bash
# row bands over NCCL: parity against the reference golden at 64x64, then the 256x256 map (config 4). Usage: gpu_rowband.sh N
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $TR bench.py --gpus $N --rowband --latent 64 --steps 10 --warmup 3 > gpurun_out/rowband_${N}_64.json 2> gpurun_out/rowband_${N}_64.err
echo "rc64=$?"; tail -3 gpurun_out/rowband_${N}_64.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/rowband_${N}_64.json").read().strip().splitlines()[-1])
    print("64x64 bands=%d ms/step %.3f e2e %.3f parity %s" % (d["n_gpus"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["parity"]))
except Exception as e:
    print("no json", e)
PY
timeout 900 $TR bench.py --gpus $N --rowband --latent 256 --steps 5 --warmup 3 --verbose > gpurun_out/rowband_${N}_256.json 2> gpurun_out/rowband_${N}_256.err
echo "rc256=$?"; tail -16 gpurun_out/rowband_${N}_256.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/rowband_${N}_256.json").read().strip().splitlines()[-1])
    print("256x256 bands=%d ms/step %.3f e2e %.3f frac %s loss %s" % (d["n_gpus"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["loss_check"]))
except Exception as e:
    print("no json", e)
PY
