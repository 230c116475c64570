#!/bin/bash
# round 2, call AH (2 GPUs): NCCL row-band tests and the driver's N = 2 bench command on the final build
O=gpurun_out/r02ah; mkdir -p $O
timeout 400 python -m pytest tests/test_rowband_gpu.py -m gpu -q --timeout 300 > $O/pytest_band.log 2>&1; echo "pytest rc=$?"
tail -3 $O/pytest_band.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02ah/bench_n2.json').read().strip().splitlines()[-1])
print("N=2: ms/step %.3f value %.1f e2e %.1f frac %.3f"%(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac']))
for k,v in d.get('other_configs',{}).items():
    print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a not in ('workload','parity','exchange')})
PY
