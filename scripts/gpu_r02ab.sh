#!/bin/bash
# round 2, call AB: bit-set constraint propagation and the sparse loss contraction: parity, then timings
O=gpurun_out/r02ab; mkdir -p $O
timeout 400 python -m pytest tests/test_propagate_gpu.py tests/test_sparse_loss_gpu.py -m gpu -q --timeout 200 -s > $O/pytest.log 2>&1; echo "pytest rc=$?"
grep -i "passed\|failed\|error\|sparse\|assert" $O/pytest.log | tail -30
timeout 200 python - <<'PY' 2>&1 | tail -20
import time, torch, numpy as np
from wavefunctiondiffusion_b200.utils.tilesets import load_wfc_mats
from wavefunctiondiffusion_b200.wfc import WFCGeneratorPriorDist
mx, my = load_wfc_mats("ice_64x64")
T, C, B, H, W = mx.shape[0], 3200, 8, 64, 64
g = torch.Generator(device="cuda").manual_seed(5)
for sharp, thr in ((6.0, 2e-3), (3.0, 1e-3), (1.0, 2e-4), (0.0, 1e-4)):
    wave = torch.softmax(sharp * torch.randn(B, H, W, C, generator=g, device="cuda"), dim=-1).to(torch.bfloat16)
    torch.cuda.synchronize(); t0 = time.time()
    gen = WFCGeneratorPriorDist.from_wave(wave, mx, my, threshold=thr)
    torch.cuda.synchronize(); t1 = time.time()
    n0 = gen.count_allowed().float().mean().item()
    seed = gen.allowed_bits.clone()
    gen.propagate_all(2048)  # warm-up: packs the connection table on the host, allocates the scratch
    gen.set_allowed_bits(seed.clone())
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); ok = gen.propagate_all(2048); b.record(); torch.cuda.synchronize()
    n1 = gen.count_allowed().float().mean().item()
    print("sharp %.0f thr %.0e: seed %.1f ms (host incl.), allowed/cell %.1f -> %.1f, ok %s, sweeps %d, propagate %.2f ms (%.3f ms/sweep)"
          % (sharp, thr, (t1 - t0) * 1e3, n0, n1, ok, gen.last_sweeps, a.elapsed_time(b), a.elapsed_time(b) / gen.last_sweeps))
PY
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > $O/bench.json 2> $O/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02ab/bench.json').read().strip().splitlines()[-1])
print("bench: ms/step %.3f value %.1f e2e %.1f roofline %.3f"%(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac']))
print(json.dumps(d['other_configs'].get('sparse_loss'), indent=1))
PY
