"""Diagnosis (not part of the library): is the groups=64 3x3 conv bound by how its activation boxes come out of HBM?
Times the same patch-schedule contraction (8 maps, 64x64, 3072 -> 3072 channels in 64 groups, N tile 48) with
  real     : the channels-last [8, 64, 64, 3072] activation tensor (each box row is 128 bytes out of a 6 KB pixel row)
  compact  : every N tile reading the SAME 64 channels of a [8, 64, 64, 64] tensor (4 MB, L2 resident, box rows contiguous)
through the test hook of the C ABI. If `compact` is not much faster, the layout of the activations is not what bounds it."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("WFD_ENABLE_TEST_HOOKS", "1")
import torch

from wavefunctiondiffusion_b200 import _native as nat


def desc(**kw):
    d = nat.TapGemmDesc()
    for k, v in kw.items():
        if k in ("tap_dx", "tap_dy"):
            arr = getattr(d, k)
            for i, t in enumerate(v):
                arr[i] = t
        else:
            setattr(d, k, v)
    return d


def timed(d, iters=20):
    lib = nat.lib()
    for _ in range(3):
        nat.check(lib.wfd_test_tapgemm(C.byref(d), 0, nat.stream_ptr()), "tapgemm")
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        nat.check(lib.wfd_test_tapgemm(C.byref(d), 0, nat.stream_ptr()), "tapgemm")
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


B, H, W, Cc, G = 8, 64, 64, 3072, 64
BN = 48
n_tiles = Cc // BN
offs = [(dy, dx) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
torch.manual_seed(0)
Bm = (torch.randn(n_tiles * BN, 9 * 64, device="cuda") * 0.05).bfloat16()
Bm.view(n_tiles * BN, 9, 64)[:, :, 48:] = 0
out = torch.zeros(B, H, W, Cc, device="cuda", dtype=torch.bfloat16)
res = {}
for name in ("real", "compact"):
    if name == "real":
        A = torch.randn(B, H, W, Cc, device="cuda").bfloat16()
        c0 = torch.arange(n_tiles, dtype=torch.int32, device="cuda") * 48
        aC = Cc
    else:
        A = torch.randn(B, H, W, 64, device="cuda").bfloat16()
        c0 = torch.zeros(n_tiles, dtype=torch.int32, device="cuda")
        aC = 64
    d = desc(a=A.data_ptr(), a_C=aC, a_W=W, a_H=H, a_B=B, a_ld=aC, b=Bm.data_ptr(), b_rows=n_tiles * BN, ldb=9 * 64, b_zstride=0,
             b_Z=1, n_taps=9, cpt=1, tap_dx=[o[1] for o in offs], tap_dy=[o[0] for o in offs], a_c0=c0.data_ptr(), BN=BN, N=Cc,
             W=W, H=H, B=B, w_box=8, h_box=16, in_stride=1, z_count=1, out=out.data_ptr(), out_fp32=0, ldc=Cc, alpha=1.0,
             dtype16=1, reuse=2)
    res[name] = timed(d)
    print("%-8s %.4f ms per launch" % (name, res[name]))
