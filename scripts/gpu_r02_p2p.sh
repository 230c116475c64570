#!/bin/bash
# round 2: peer-mailbox transport of the row bands - parity worker on N GPUs (P2P and NCCL), then the 256x256 strong-scaling
# bench with both transports. Usage: gpu_r02_p2p.sh N
N=${1:-2}
O=gpurun_out/r02p2p; mkdir -p $O
for p2p in 1 0; do
  WFD_BAND_P2P=$p2p timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$p2p tests/band_nccl_worker.py > $O/worker_n${N}_p2p$p2p.log 2>&1
  echo "worker N=$N p2p=$p2p rc=$?"; grep "NCCL_BAND_OK\|graph=\|Error\|error\|assert" $O/worker_n${N}_p2p$p2p.log | head -6
done
for p2p in 1 0; do
  WFD_VERBOSE=1 WFD_BAND_P2P=$p2p timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$p2p bench.py --gpus $N --rowband --latent 256 --steps 10 --warmup 3 --no-cpu-baseline > $O/rowband_n${N}_p2p$p2p.json 2> $O/rowband_n${N}_p2p$p2p.err
  echo "rowband N=$N p2p=$p2p rc=$?"
  python - $O/rowband_n${N}_p2p$p2p.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  ms/step %.3f value %.2f e2e %.2f loss %s"%(d['ms_per_step'], d['value'], d['e2e']['value'], d.get('loss_check')))
except Exception as e:
    print("  FAILED", e)
PY
done
grep -h "peer mailboxes" $O/rowband_n${N}_p2p1.err | head -2
