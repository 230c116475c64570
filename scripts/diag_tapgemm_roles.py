"""Where does a tapgemm launch spend its time? Cycle counters of CTA 0 per warp role for decoder-like conv shapes."""
import os; os.environ.setdefault("WFD_ENABLE_TEST_HOOKS", "1")
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from wavefunctiondiffusion_b200 import _native as nat
from test_tapgemm_gpu import _desc

def run(cin, cout, groups, BN, Bn=8, H=64, W=64, taps9=True, reuse=0, cg=0):
    cin_g, cout_g = cin // groups, cout // groups
    n_tiles = cout // BN
    gpt = max(1, BN // cout_g)                      # groups per N tile
    cpt = (gpt * cin_g + 63) // 64 if cout_g <= BN else (cin_g + 63) // 64
    ntap = 9 if taps9 else 1
    x = torch.randn(Bn, H, W, cin, device="cuda").bfloat16()
    Bm = torch.randn(n_tiles * BN, ntap * cpt * 64, device="cuda").bfloat16()
    c0 = torch.tensor([(t * BN // cout_g) * cin_g for t in range(n_tiles)], dtype=torch.int32, device="cuda")
    out = torch.zeros(Bn, H, W, cout, device="cuda", dtype=torch.bfloat16)
    bias = torch.randn(cout, device="cuda")
    dbg = torch.zeros(16, dtype=torch.int64, device="cuda")
    taps = [(ky - 1, kx - 1) for ky in range(3) for kx in range(3)] if taps9 else [(0, 0)]
    d = _desc(a=x.data_ptr(), a_C=cin, a_W=W, a_H=H, a_B=Bn, a_ld=cin, b=Bm.data_ptr(), b_rows=Bm.shape[0],
              ldb=Bm.shape[1], b_zstride=0, b_Z=1, n_taps=ntap, cpt=cpt, tap_dx=[t[1] for t in taps],
              tap_dy=[t[0] for t in taps], a_c0=c0.data_ptr(), BN=BN, N=cout, W=W, H=H, B=Bn, w_box=64, h_box=2,
              in_stride=1, z_count=1, out=out.data_ptr(), out_fp32=0, ldc=cout, bias=bias.data_ptr(), alpha=1.0,
              dtype16=1, dbg=dbg.data_ptr(), reuse=reuse, cg=cg)
    lib = nat.lib()
    for _ in range(2):
        nat.check(lib.wfd_test_tapgemm(C.byref(d), 0, nat.stream_ptr()))
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        nat.check(lib.wfd_test_tapgemm(C.byref(d), 0, nat.stream_ptr()))
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    v = dbg.tolist()
    tiles = (Bn * H * W // 128) * n_tiles
    per_cta = tiles / 148
    flops = 2.0 * Bn * H * W * cout * cin_g * ntap
    print("cg %d " % cg + "cin %d cout %d g %d BN %d cpt %d taps %d: %.3f ms  %.0f TF/s | tiles/CTA %.1f | producer total %d wait_empty %d | mma total %d wait_tempty %d wait_full %d | epi total %d wait_tfull %d busy %d reuse %d | Bprod total %d wait %d | issueA %d issueB %d (cycles; per tile: total %.0f)"
          % (cin, cout, groups, BN, cpt, ntap, ms, flops / ms / 1e9, per_cta, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], reuse, v[8], v[9], v[10], v[11], v[2] / per_cta))

for cg in (1, 2):
    run(384, 384, 1, 192, cg=cg)
    run(768, 768, 4, 192, cg=cg)
    run(384, 384, 1, 128, cg=cg)
