"""Plain GEMM with the dense-conv problem size vs the conv itself: is the activation box pattern the limiter?"""
import os; os.environ.setdefault("WFD_ENABLE_TEST_HOOKS", "1")
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from wavefunctiondiffusion_b200 import _native as nat
from test_tapgemm_gpu import _desc

def timeit(d, flops, label):
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
    print("%-60s %.3f ms %.0f TF/s" % (label, ms, flops / ms / 1e9))

for cg in (1, 2):
    for (M, K, N, BN) in [(32768, 3456, 384, 192), (32768, 3456, 768, 256), (32768, 3456, 384, 128), (32768, 6912, 768, 192)]:
        A = torch.randn(M, K, device="cuda").bfloat16()
        Bm = torch.randn(((N + BN - 1) // BN) * BN, K, device="cuda").bfloat16()
        out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
        d = _desc(a=A.data_ptr(), a_C=K, a_W=M, a_H=1, a_B=1, a_ld=K, b=Bm.data_ptr(), b_rows=Bm.shape[0], ldb=K, b_zstride=0,
                  b_Z=1, n_taps=1, cpt=K // 64, tap_dx=[0], tap_dy=[0], a_c0=None, BN=BN, N=N, W=M, H=1, B=1, w_box=128,
                  h_box=1, in_stride=1, z_count=1, out=out.data_ptr(), out_fp32=0, ldc=N, alpha=1.0, dtype16=1, cg=cg)
        timeit(d, 2.0 * M * N * K, "plain cg=%d M=%d K=%d N=%d BN=%d" % (cg, M, K, N, BN))
    # the conv with the same arithmetic: 8 x 64x64 x 384 -> 384, 9 taps
    x = torch.randn(8, 64, 64, 384, device="cuda").bfloat16()
    Bm = torch.randn(384, 9 * 6 * 64, device="cuda").bfloat16()
    out = torch.zeros(8, 64, 64, 384, device="cuda", dtype=torch.bfloat16)
    taps = [(ky - 1, kx - 1) for ky in range(3) for kx in range(3)]
    for shifted in (1, 0):
        d = _desc(a=x.data_ptr(), a_C=384, a_W=64, a_H=64, a_B=8, a_ld=384, b=Bm.data_ptr(), b_rows=384, ldb=Bm.shape[1],
                  b_zstride=0, b_Z=1, n_taps=9, cpt=6, tap_dx=[t[1] * shifted for t in taps], tap_dy=[t[0] * shifted for t in taps],
                  a_c0=None, BN=192, N=384, W=64, H=64, B=8, w_box=64, h_box=2, in_stride=1, z_count=1, out=out.data_ptr(),
                  out_fp32=0, ldc=384, alpha=1.0, dtype16=1, cg=cg)
        timeit(d, 2.0 * 32768 * 384 * 3456, "conv cg=%d shifted=%d" % (cg, shifted))
