"""Is the k-iteration rate of the big GEMMs limited per SM or chip-wide? Same GEMM on 148 / 74 / 37 CTAs."""
import os; os.environ.setdefault("WFD_ENABLE_TEST_HOOKS", "1")
import ctypes as C, os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from wavefunctiondiffusion_b200 import _native as nat
from test_tapgemm_gpu import _desc

M, K, N, BN = 32768, 3200 * 2, 3200, 256
A = torch.randn(M, K, device="cuda").bfloat16()
Bm = torch.randn(3328, K, device="cuda").bfloat16()
out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
dbg = torch.zeros(16, dtype=torch.int64, device="cuda")
d = _desc(a=A.data_ptr(), a_C=K, a_W=M, a_H=1, a_B=1, a_ld=K, b=Bm.data_ptr(), b_rows=3328, ldb=K, b_zstride=0, b_Z=1,
          n_taps=1, cpt=K // 64, tap_dx=[0], tap_dy=[0], a_c0=None, BN=BN, N=N, W=M, H=1, B=1, w_box=128, h_box=1,
          in_stride=1, z_count=1, out=out.data_ptr(), out_fp32=0, ldc=N, alpha=1.0, dtype16=1, dbg=dbg.data_ptr(), n_group=int(os.environ.get('NG', '0')))
lib = nat.lib()
for _ in range(2):
    nat.check(lib.wfd_test_tapgemm(C.byref(d), 0, nat.stream_ptr()))
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    nat.check(lib.wfd_test_tapgemm(C.byref(d), 0, nat.stream_ptr()))
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
v = dbg.tolist()
grid = int(os.environ.get("WFD_GRID_LIMIT", "148") or 148)
tiles = (M // 128) * 13
kiters = tiles / min(grid, 148) * (K // 64)
print("n_group " + os.environ.get("NG", "0") + " grid %s: %.3f ms %.0f TF/s | cycles per k-iter %.0f | producer wait_empty %.0f%% mma wait_full %.0f%%"
      % (os.environ.get("WFD_GRID_LIMIT", "148"), ms, 2.0 * M * N * K / ms / 1e9, v[2] / kiters, 100.0 * v[1] / max(v[0], 1), 100.0 * v[4] / max(v[2], 1)))
