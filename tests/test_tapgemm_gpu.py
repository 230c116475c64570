"""tapgemm (tcgen05/TMA/TMEM contraction kernel) against torch fp32 on the same rounded operands.

The kernel is reached through the C ABI test hook wfd_test_tapgemm; use_ref=1 runs the CUDA-core checking kernel on the
same descriptors, which separates descriptor/packing mistakes from tensor-core mistakes.
"""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from wavefunctiondiffusion_b200 import _native as nat

pytestmark = pytest.mark.gpu


def _desc(**kw):
    d = nat.TapGemmDesc()
    for k, v in kw.items():
        if k in ("tap_dx", "tap_dy"):
            arr = getattr(d, k)
            for i, t in enumerate(v):
                arr[i] = t
        else:
            setattr(d, k, v)
    return d


def _run(d, use_ref):
    nat.check(nat.lib().wfd_test_tapgemm(C.byref(d), use_ref, nat.stream_ptr()), "wfd_test_tapgemm")
    torch.cuda.synchronize()


def _dt(bf16):
    return torch.bfloat16 if bf16 else torch.float16


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-30)).item()


@pytest.mark.parametrize("use_ref", [1, 0])
@pytest.mark.parametrize("bf16", [1, 0])
@pytest.mark.parametrize("M,K,N,BN", [(256, 128, 64, 64), (1024, 384, 384, 192), (512, 64, 48, 48), (384, 1152, 256, 256)])
def test_plain_gemm(M, K, N, BN, bf16, use_ref):
    torch.manual_seed(0)
    dt = _dt(bf16)
    A = torch.randn(M, K, device="cuda").to(dt)
    Bm = torch.randn(N, K, device="cuda").to(dt)
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda").to(dt)
    out = torch.zeros(M, N, device="cuda", dtype=dt)
    d = _desc(a=A.data_ptr(), a_C=K, a_W=M, a_H=1, a_B=1, a_ld=K, b=Bm.data_ptr(), b_rows=N, ldb=K, b_zstride=0, b_Z=1,
              n_taps=1, cpt=K // 64, tap_dx=[0], tap_dy=[0], a_c0=None, BN=BN, N=N, W=M, H=1, B=1, w_box=128, h_box=1,
              in_stride=1, z_count=1, out=out.data_ptr(), out_fp32=0, ldc=N, bias=bias.data_ptr(),
              residual=res.data_ptr(), alpha=0.5, dtype16=bf16)
    _run(d, use_ref)
    want = 0.5 * (A.float() @ Bm.float().t()) + bias + res.float()
    assert _rel(out, want) < (6e-3 if bf16 else 1e-3)


@pytest.mark.parametrize("use_ref", [1, 0])
def test_fp32_out_rowdot_batched(use_ref):
    """scores-style: per-sample B operand, fp32 output, plus the row-dot epilogue used by the WFC contraction."""
    torch.manual_seed(1)
    Bz, M, K, N = 3, 256, 128, 256
    A = torch.randn(Bz, M, K, device="cuda").bfloat16()
    Bm = torch.randn(Bz, N, K, device="cuda").bfloat16()
    dots = torch.randn(Bz, M, N, device="cuda").bfloat16()
    out = torch.zeros(Bz, M, N, device="cuda")
    rowdot = torch.zeros(Bz * M, device="cuda", dtype=torch.int64)  # fixed point 2^40 (order-independent atomics)
    d = _desc(a=A.data_ptr(), a_C=K, a_W=M, a_H=1, a_B=Bz, a_ld=K, b=Bm.data_ptr(), b_rows=N, ldb=K, b_zstride=N * K,
              b_Z=Bz, n_taps=1, cpt=K // 64, tap_dx=[0], tap_dy=[0], BN=128, N=N, W=M, H=1, B=Bz, w_box=128, h_box=1,
              in_stride=1, z_count=1, b_bbmul=1, out=out.data_ptr(), out_fp32=1, ldc=N, alpha=1.0,
              rowdot=rowdot.data_ptr(), dotsrc=dots.data_ptr(), ld_dot=N, dtype16=1)
    _run(d, use_ref)
    want = torch.bmm(A.float(), Bm.float().transpose(1, 2))
    assert _rel(out, want) < 1e-4
    assert _rel(rowdot.double().view(Bz, M) / 2.0 ** 40, (want * dots.float()).sum(-1)) < 1e-3


@pytest.mark.parametrize("use_ref", [1, 0])
def test_shared_a_z_batch(use_ref):
    """v^T-style: A (weights) shared by every z, B operand and output indexed by z."""
    torch.manual_seed(2)
    Z, M, K, N = 2, 384, 384, 512
    A = torch.randn(M, K, device="cuda").bfloat16()
    Bm = torch.randn(Z, N, K, device="cuda").bfloat16()
    out = torch.zeros(Z, M, N, device="cuda", dtype=torch.bfloat16)
    d = _desc(a=A.data_ptr(), a_C=K, a_W=M, a_H=1, a_B=1, a_ld=K, b=Bm.data_ptr(), b_rows=N, ldb=K, b_zstride=N * K,
              b_Z=Z, n_taps=1, cpt=K // 64, tap_dx=[0], tap_dy=[0], BN=256, N=N, W=M, H=1, B=1, w_box=128, h_box=1,
              in_stride=1, z_count=Z, b_zmul=1, out=out.data_ptr(), out_fp32=0, ldc=N, c_zstride=M * N, alpha=1.0,
              dtype16=1)
    _run(d, use_ref)
    want = torch.einsum("mk,znk->zmn", A.float(), Bm.float())
    assert _rel(out, want) < 6e-3


def _pack_conv(w, groups, BN, bwd=False, reuse=False):
    """Python restatement of the C++ packer for the test: returns (Bm [n_tiles*BN, taps*cpt*64], a_c0, cpt, N).
    reuse=1 stores the 64-wide chunks in the (dx, cc, dy) order of the row-reuse schedule, reuse=2 in the (cc, dy, dx)
    order of the patch schedule."""
    cout, cin_g, k, _ = w.shape
    cin = cin_g * groups
    cout_g = cout // groups
    N = cin if bwd else cout
    rpg = cin_g if bwd else cout_g
    kpg = cout_g if bwd else cin_g
    n_tiles = (N + BN - 1) // BN
    c0, cpt = [], 1
    for t in range(n_tiles):
        n0, n1 = t * BN, min(N, (t + 1) * BN)
        g_lo, g_hi = n0 // rpg, (n1 - 1) // rpg
        c0.append(g_lo * kpg)
        cpt = max(cpt, ((g_hi + 1) * kpg - c0[-1] + 63) // 64)
    taps = k * k
    Bm = torch.zeros(n_tiles * BN, taps * cpt * 64)
    wf = w.reshape(cout, cin_g, taps)
    sign = -1 if bwd else 1
    offs = [(sign * (ky - k // 2), sign * (kx - k // 2)) for ky in range(k) for kx in range(k)]  # (dy, dx) per tap
    for n in range(N):
        g, t = n // rpg, n // BN
        for kk in range(kpg):
            pos = g * kpg + kk - c0[t]
            cc, within = pos // 64, pos % 64
            if not bwd:
                vals = wf[n, kk, :]
            else:
                vals = wf[g * cout_g + kk, n - g * cin_g, :]
            if reuse == 2:
                chunk = torch.tensor([cc * 9 + (dy + 1) * 3 + (dx + 1) for dy, dx in offs])
            elif reuse:
                chunk = torch.tensor([((dx + 1) * cpt + cc) * 3 + (dy + 1) for dy, dx in offs])
            else:
                chunk = torch.arange(taps) * cpt + cc
            Bm[n, chunk * 64 + within] = vals
    return Bm, torch.tensor(c0, dtype=torch.int32), cpt, N


@pytest.mark.parametrize("reuse", [0, 1, 2])
@pytest.mark.parametrize("use_ref", [1, 0])
@pytest.mark.parametrize("cin,cout,groups,BN,H,W", [(64, 64, 1, 64, 16, 16), (192, 96, 4, 48, 8, 64), (96, 192, 2, 96, 32, 32),
                                                   (384, 384, 1, 192, 64, 64), (3072, 3072, 64, 48, 32, 16)])
def test_conv3x3_forward_and_backward_data(cin, cout, groups, BN, H, W, use_ref, reuse):
    if reuse == 1 and BN > 96:
        pytest.skip("row-reuse schedule is used for tiles that leave >= 3 stages")
    if reuse == 2 and (H % 16 or W % 8):
        pytest.skip("the patch schedule tiles the plane into 16 x 8 pixel patches")
    torch.manual_seed(3)
    Bn = 2
    x = torch.randn(Bn, cin, H, W, device="cuda").bfloat16()
    w = (torch.randn(cout, cin // groups, 3, 3) * 0.1).bfloat16().float()
    bias = torch.randn(cout, device="cuda")
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    w_box = min(W, 128)
    h_box = 128 // w_box
    if reuse == 2:
        w_box, h_box = 8, 16
    # forward
    Bm, c0, cpt, N = _pack_conv(w, groups, BN, reuse=reuse)
    Bm = Bm.cuda().bfloat16()
    c0 = c0.cuda()
    out = torch.zeros(Bn, H, W, cout, device="cuda", dtype=torch.bfloat16)
    gn = torch.zeros(Bn * 8 * 2, device="cuda", dtype=torch.int64)  # fixed point 2^20
    taps = [(ky - 1, kx - 1) for ky in range(3) for kx in range(3)]
    d = _desc(a=x_nhwc.data_ptr(), a_C=cin, a_W=W, a_H=H, a_B=Bn, a_ld=cin, b=Bm.data_ptr(), b_rows=Bm.shape[0],
              ldb=Bm.shape[1], b_zstride=0, b_Z=1, n_taps=9, cpt=cpt, tap_dx=[t[1] for t in taps],
              tap_dy=[t[0] for t in taps], a_c0=c0.data_ptr(), BN=BN, N=N, W=W, H=H, B=Bn, w_box=w_box, h_box=h_box,
              in_stride=1, z_count=1, out=out.data_ptr(), out_fp32=0, ldc=cout, bias=bias.data_ptr(), alpha=1.0,
              gn_sums=gn.data_ptr(), gn_cpg=cout // 8, dtype16=1, reuse=reuse)
    _run(d, use_ref)
    want = F.conv2d(x.float(), w.cuda(), bias, padding=1, groups=groups)
    got = out.permute(0, 3, 1, 2).float()
    assert _rel(got, want) < 6e-3
    # fused GroupNorm statistics of the stored (rounded) output
    # (the tensor-core epilogue reduces 16-column chunks, so it needs channels-per-group % 16 == 0; the decoder only
    # fuses in that case and runs the standalone statistics kernel otherwise)
    if use_ref or (cout // 8) % 16 == 0:
        g8 = got.reshape(Bn, 8, -1)
        gnf = gn.double().view(Bn, 8, 2) / 2.0 ** 20
        assert (gnf[..., 0] - g8.sum(-1)).abs().max() < 2e-3 * g8.abs().sum(-1).max()
        assert _rel(gnf[..., 1], (g8 * g8).sum(-1)) < 2e-3
    # backward data: dX = conv_transpose(dY, w)
    dy = torch.randn(Bn, cout, H, W, device="cuda").bfloat16()
    dy_nhwc = dy.permute(0, 2, 3, 1).contiguous()
    BNb = BN if (cin // groups) % 16 == 0 and (cin // groups) <= 256 and BN % (cin // groups) == 0 else 64
    Bm2, c02, cpt2, N2 = _pack_conv(w, groups, BNb, bwd=True, reuse=reuse)
    Bm2 = Bm2.cuda().bfloat16()
    c02 = c02.cuda()
    dx = torch.zeros(Bn, H, W, cin, device="cuda", dtype=torch.bfloat16)
    d = _desc(a=dy_nhwc.data_ptr(), a_C=cout, a_W=W, a_H=H, a_B=Bn, a_ld=cout, b=Bm2.data_ptr(), b_rows=Bm2.shape[0],
              ldb=Bm2.shape[1], b_zstride=0, b_Z=1, n_taps=9, cpt=cpt2, tap_dx=[-t[1] for t in taps],
              tap_dy=[-t[0] for t in taps], a_c0=c02.data_ptr(), BN=BNb, N=N2, W=W, H=H, B=Bn, w_box=w_box, h_box=h_box,
              in_stride=1, z_count=1, out=dx.data_ptr(), out_fp32=0, ldc=cin, alpha=1.0, dtype16=1, reuse=reuse)
    _run(d, use_ref)
    want = F.conv_transpose2d(dy.float(), w.cuda(), padding=1, groups=groups)
    assert _rel(dx.permute(0, 3, 1, 2), want) < 6e-3


@pytest.mark.parametrize("use_ref", [1, 0])
def test_four_tap_neighbour_contraction(use_ref):
    """WFC-style: 4 taps (right, left, below, above) with zero fill at the map border."""
    torch.manual_seed(4)
    Bn, H, W, T = 2, 16, 16, 128
    p = torch.rand(Bn, H, W, T, device="cuda").bfloat16()
    mats = (torch.rand(4, T, T, device="cuda") < 0.05).bfloat16()  # [tap][i][j]
    Bm = torch.cat([mats[t] for t in range(4)], dim=1).contiguous()  # [T, 4T]
    out = torch.zeros(Bn, H, W, T, device="cuda", dtype=torch.bfloat16)
    d = _desc(a=p.data_ptr(), a_C=T, a_W=W, a_H=H, a_B=Bn, a_ld=T, b=Bm.data_ptr(), b_rows=T, ldb=4 * T, b_zstride=0,
              b_Z=1, n_taps=4, cpt=T // 64, tap_dx=[1, -1, 0, 0], tap_dy=[0, 0, 1, -1], BN=128, N=T, W=W, H=H, B=Bn,
              w_box=16, h_box=8, in_stride=1, z_count=1, out=out.data_ptr(), out_fp32=0, ldc=T, alpha=1.0, dtype16=1)
    _run(d, use_ref)
    pf = p.float()
    want = torch.zeros(Bn, H, W, T, device="cuda")
    want[:, :, :-1] += torch.einsum("ij,bhwj->bhwi", mats[0].float(), pf[:, :, 1:])
    want[:, :, 1:] += torch.einsum("ij,bhwj->bhwi", mats[1].float(), pf[:, :, :-1])
    want[:, :-1] += torch.einsum("ij,bhwj->bhwi", mats[2].float(), pf[:, 1:])
    want[:, 1:] += torch.einsum("ij,bhwj->bhwi", mats[3].float(), pf[:, :-1])
    assert _rel(out, want) < 6e-3


@pytest.mark.parametrize("use_ref", [1, 0])
@pytest.mark.parametrize("M,K,N,BN,Bn", [(256, 128, 64, 64, 1), (512, 256, 384, 192, 2), (1024, 192, 48, 48, 1)])
def test_m_major_a_operand(M, K, N, BN, Bn, use_ref):
    """A handed over transposed in memory ([K][M], M contiguous) - the attention backward products dV = P^T dO and
    dK = dS^T Q read P / dS as they are stored instead of going through a transpose pass."""
    torch.manual_seed(7)
    At = torch.randn(Bn, K, M, device="cuda").bfloat16()          # [sample][k][m]
    Bm = torch.randn(Bn, N, K, device="cuda").bfloat16()
    out = torch.zeros(Bn, M, N, device="cuda", dtype=torch.bfloat16)
    d = _desc(a=At.data_ptr(), a_C=K, a_W=M, a_H=1, a_B=Bn, a_ld=M, b=Bm.data_ptr(), b_rows=N, ldb=K, b_zstride=N * K,
              b_Z=Bn, n_taps=1, cpt=K // 64, tap_dx=[0], tap_dy=[0], a_c0=None, BN=BN, N=N, W=M, H=1, B=Bn, w_box=128,
              h_box=1, in_stride=1, z_count=1, b_bbmul=1, out=out.data_ptr(), out_fp32=0, ldc=N, alpha=1.0, dtype16=1,
              a_mn_major=1)
    _run(d, use_ref)
    want = torch.einsum("bkm,bnk->bmn", At.float(), Bm.float())
    assert _rel(out, want) < 6e-3


@pytest.mark.parametrize("n_group", [0, 4])
def test_block_sparse_skip_list_is_bit_identical(n_group):
    """The WFC contraction visits only the K chunks of the adjacency operand that hold an allowed pair (wfc.cu): the
    skipped chunks are all zero, so the result must equal the dense walk bit for bit (and the CUDA-core checker)."""
    torch.manual_seed(6)
    Bn, H, W, T, BN = 4, 32, 32, 512, 256
    p = torch.rand(Bn, H, W, T, device="cuda").bfloat16()
    mats = (torch.rand(4, T, T, device="cuda") < 0.02).bfloat16()
    keep = torch.rand(4, T // BN, T // 64, device="cuda") < 0.4          # [tap][N tile][chunk] blocks that survive
    keep[:, 1, :] = False
    keep[0, 1, 3] = False                                                  # N tile 1 ends up with no chunk at all
    mask = keep.repeat_interleave(BN, 1).repeat_interleave(64, 2).to(mats.dtype)
    mats = mats * mask
    Bm = torch.cat([mats[t] for t in range(4)], dim=1).contiguous()      # [T, 4T]
    cpt, k_iters, n_tiles = T // 64, 4 * (T // 64), T // BN
    tdx, tdy = [1, -1, 0, 0], [0, 0, 1, -1]
    lst = torch.zeros(n_tiles, k_iters, dtype=torch.int32)
    cnt = torch.zeros(n_tiles, dtype=torch.int32)
    for n in range(n_tiles):
        c = 0
        for kc in range(k_iters):
            tap, cc = divmod(kc, cpt)
            if bool(Bm[n * BN:(n + 1) * BN, kc * 64:(kc + 1) * 64].any()):
                lst[n, c] = (tdx[tap] + 1) | ((tdy[tap] + 1) << 2) | (cc << 4) | (kc << 16)
                c += 1
        cnt[n] = max(c, 1)
    assert cnt.min().item() == 1 and cnt.sum().item() < n_tiles * k_iters
    lst, cnt = lst.cuda(), cnt.cuda()
    outs = []
    for use_list, use_ref in ((True, 0), (False, 0), (False, 1)):
        out = torch.zeros(Bn, H, W, T, device="cuda", dtype=torch.bfloat16)
        d = _desc(a=p.data_ptr(), a_C=T, a_W=W, a_H=H, a_B=Bn, a_ld=T, b=Bm.data_ptr(), b_rows=T, ldb=4 * T, b_zstride=0,
                  b_Z=1, n_taps=4, cpt=cpt, tap_dx=tdx, tap_dy=tdy, BN=BN, N=T, W=W, H=H, B=Bn, w_box=32, h_box=4,
                  in_stride=1, z_count=1, out=out.data_ptr(), out_fp32=0, ldc=T, alpha=1.0, dtype16=1, n_group=n_group,
                  kc_list=lst.data_ptr() if use_list else None, kc_cnt=cnt.data_ptr() if use_list else None)
        _run(d, use_ref)
        outs.append(out)
    assert torch.equal(outs[0], outs[1])
    assert _rel(outs[0], outs[2]) < 6e-3


@pytest.mark.parametrize("use_ref", [1, 0])
@pytest.mark.parametrize("n_group", [0, 4])
def test_persistent_many_tiles_per_cta(use_ref, n_group):
    """More tiles than SMs: every CTA loops over several tiles, exercising the TMEM double buffer and the smem ring
    across tile boundaries (M = 128 * 450 rows, 3 N tiles -> 1350 tiles on 148 SMs)."""
    torch.manual_seed(5)
    M, K, N, BN = 128 * 450, 192, 144, 48
    A = torch.randn(M, K, device="cuda").bfloat16()
    Bm = torch.randn(N, K, device="cuda").bfloat16()
    bias = torch.randn(N, device="cuda")
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    d = _desc(a=A.data_ptr(), a_C=K, a_W=M, a_H=1, a_B=1, a_ld=K, b=Bm.data_ptr(), b_rows=N, ldb=K, b_zstride=0, b_Z=1,
              n_taps=1, cpt=K // 64, tap_dx=[0], tap_dy=[0], a_c0=None, BN=BN, N=N, W=M, H=1, B=1, w_box=128, h_box=1,
              in_stride=1, z_count=1, out=out.data_ptr(), out_fp32=0, ldc=N, bias=bias.data_ptr(), alpha=1.0, dtype16=1,
              n_group=n_group)
    _run(d, use_ref)
    want = A.float() @ Bm.float().t() + bias
    assert _rel(out, want) < 6e-3
