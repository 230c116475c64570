#!/bin/bash
# round 2, call AK: constraint propagation with one warp per cell (default) against one CTA per cell (WFD_PROP_CTA=1)
O=gpurun_out/r02ak; mkdir -p $O
timeout 300 python -m pytest tests/test_propagate_gpu.py -m gpu -q --timeout 200 > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
for MODE in 0 1; do
echo "WFD_PROP_CTA=$MODE"
WFD_PROP_CTA=$MODE timeout 200 python - <<'PY' 2>&1 | tail -6
import torch
from wavefunctiondiffusion_b200.utils.tilesets import load_wfc_mats
from wavefunctiondiffusion_b200.wfc import WFCGeneratorPriorDist
mx, my = load_wfc_mats("ice_64x64")
T, C, B, H, W = mx.shape[0], 3200, 8, 64, 64
g = torch.Generator(device="cuda").manual_seed(5)
for sharp, thr in ((6.0, 2e-3), (3.0, 1e-3), (1.0, 2e-4), (0.0, 1e-4)):
    wave = torch.softmax(sharp * torch.randn(B, H, W, C, generator=g, device="cuda"), dim=-1).to(torch.bfloat16)
    gen = WFCGeneratorPriorDist.from_wave(wave, mx, my, threshold=thr)
    n0 = gen.count_allowed().float().mean().item()
    seed = gen.allowed_bits.clone()
    gen.propagate_all(2048)
    best = 1e9
    for _ in range(3):
        gen.set_allowed_bits(seed.clone())
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); ok = gen.propagate_all(2048); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print("sharp %.0f thr %.0e: allowed/cell %.1f -> %.1f, ok %s, sweeps %d, propagate %.2f ms (%.3f ms/sweep), checksum %d"
          % (sharp, thr, n0, gen.count_allowed().float().mean().item(), ok, gen.last_sweeps, best, best / gen.last_sweeps, int(gen.allowed_bits.long().sum().item())))
PY
done
