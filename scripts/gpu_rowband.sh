#!/bin/bash
# row bands over NCCL: parity against the reference golden at 64x64, then the 256x256 map (config 4).
# Usage: gpu_rowband.sh N [sizes]   (tight timeouts: a hang must not eat the GPU budget)
cd "$(dirname "$0")/.."
N=${1:-2}
SIZES=${2:-"64 256"}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
for L in $SIZES; do
  T0=$(date +%s)
  timeout -k 5 ${ROWBAND_TIMEOUT:-170} $TR bench.py --gpus $N --rowband --latent $L --steps 10 --warmup 3 --verbose ${ROWBAND_FLAGS} > gpurun_out/rowband_${N}_${L}.json 2> gpurun_out/rowband_${N}_${L}.err
  echo "latent $L rc=$? wall $(( $(date +%s) - T0 )) s"
  grep -E "tapgemm|gn_|attn|transpose|profiled" gpurun_out/rowband_${N}_${L}.err | tail -8
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/rowband_${N}_${L}.json").read().strip().splitlines()[-1])
    print("${L}x${L} bands=%d graph=%s ms/step %.3f e2e %.3f frac %.3f loss %s parity %s" % (d["n_gpus"], d.get("cuda_graph"), d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"] or 0, d["loss_check"], d["parity"]))
except Exception as e:
    print("no json", e)
PY
done
