"""GPU parity of the CUDA path (through the C ABI, via the host-side mirrors) against the golden vectors of the
unmodified reference and against the CPU oracle.

Tolerances (bf16 operands, fp32 accumulation; stated against the fp64 reference values):
  logits  rel-L2 <= 2e-2        loss  rel <= 1e-3        latent gradient  rel-L2 <= 5e-2
  argmax  bit-exact on identical waves
"""
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch

from oracle import wfd_oracle as orc
from wavefunctiondiffusion_b200 import _native as nat
from wavefunctiondiffusion_b200.utils.tilesets import load_wfc_mats, round_up_channels
from wavefunctiondiffusion_b200.wf_diffusers.native_ops import GuidanceLossFunction, argmax_tiles
from wavefunctiondiffusion_b200.wf_diffusers.wf_vae import AutoencoderTile
from wavefunctiondiffusion_b200.wfdf.losses import WFCLossEinsum

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL_LOGITS, TOL_LOSS, TOL_GRAD = 2e-2, 1e-3, 5e-2


def _rel(a, b):
    a = torch.as_tensor(np.asarray(a.detach() if torch.is_tensor(a) else a), dtype=torch.float64)
    b = torch.as_tensor(np.asarray(b.detach() if torch.is_tensor(b) else b), dtype=torch.float64)
    return ((a - b).norm() / b.norm()).item()


def _load_small(name):
    z = np.load(os.path.join(GOLD, name))
    cfg = json.loads(str(z["cfg_json"]))
    T = int(z["T"])
    mats = np.unpackbits(z["mats"])[: 2 * T * T].reshape(2, T, T).astype(bool)
    torch.manual_seed(int(z["seed"]))
    net = AutoencoderTile(**cfg).eval().cuda()
    return z, cfg, mats, net


@pytest.fixture(autouse=True)
def _reset_debug():
    yield
    nat.lib().wfd_debug_set_ref_gemm(0)


@pytest.mark.parametrize("ref_gemm", [1, 0])
@pytest.mark.parametrize("name", ["tiny_step.npz", "tiny_down_step.npz"])
def test_decode_logits_small(name, ref_gemm):
    z, cfg, mats, net = _load_small(name)
    nat.lib().wfd_debug_set_ref_gemm(ref_gemm)
    lat = torch.from_numpy(z["latents"]).cuda()
    with torch.no_grad():
        logits = net.decode(lat / 0.18215).sample
    assert logits.shape == z["logits_f64"].shape and logits.dtype == torch.float32
    assert _rel(logits.cpu(), z["logits_f64"]) < TOL_LOGITS


@pytest.mark.parametrize("ref_gemm", [1, 0])
def test_guidance_block_as_the_reference_pipeline_writes_it(ref_gemm):
    """decode -> softmax(axis=1) -> WFCLossEinsum -> autograd.grad, i.e. pipeline_wfd.py:609-613 line by line."""
    z, cfg, mats, net = _load_small("tiny_step.npz")
    nat.lib().wfd_debug_set_ref_gemm(ref_gemm)
    wfc_loss = WFCLossEinsum(mats[0], mats[1]).cuda()
    latents = torch.from_numpy(z["latents"]).cuda()
    with torch.enable_grad():
        _latents = latents.detach().requires_grad_()
        image = net.decode(1 / 0.18215 * _latents).sample
        wave = image.softmax(axis=1)
        _loss = wfc_loss(wave) * 5.0
        (grad,) = torch.autograd.grad(_loss.sum(), _latents)
    assert _rel(_loss.cpu(), z["loss_f64"]) < TOL_LOSS
    assert _rel(grad.cpu(), z["grad_f64"]) < TOL_GRAD


@pytest.mark.parametrize("name", ["tiny_step.npz", "tiny_down_step.npz"])
@pytest.mark.parametrize("ref_gemm", [1, 0])
def test_fused_guidance_step_small(ref_gemm, name):
    z, cfg, mats, net = _load_small(name)
    nat.lib().wfd_debug_set_ref_gemm(ref_gemm)
    wfc_loss = WFCLossEinsum(mats[0], mats[1]).cuda()
    lat = torch.from_numpy(z["latents"]).cuda().requires_grad_()
    loss = GuidanceLossFunction.apply(lat, net._native, wfc_loss.native_handle(cfg["out_channels"], lat.device),
                                      1.0 / 0.18215, 5.0)
    (grad,) = torch.autograd.grad(loss.sum(), lat)
    assert _rel(loss.cpu(), z["loss_f64"]) < TOL_LOSS
    assert _rel(grad.cpu(), z["grad_f64"]) < TOL_GRAD
    # softmax=True entry of the loss module on the logits gives the same loss
    with torch.no_grad():
        logits = net.decode(lat.detach() / 0.18215).sample
        l2 = wfc_loss(logits, softmax=True) * 5.0
    assert _rel(l2.cpu(), z["loss_f64"]) < TOL_LOSS


@pytest.mark.parametrize("name", ["tiny_step.npz", "ice_c1_step.npz"])
def test_groupnorm_apply_inside_the_conv_matches_the_standalone_pass(name, monkeypatch):
    """GroupNorm+SiLU in front of a 3x3 conv runs in the conv's operand path (transform warps rewrite the landed
    activation box in shared memory, reference resnet.py:461-462, 483-489). Same arithmetic, same rounding points as the
    standalone gn_apply pass (WFD_FUSE_APPLY=0; 2 = every patch-schedule conv): the logits agree bit for bit."""
    z = np.load(os.path.join(GOLD, name))
    cfg = json.loads(str(z["cfg_json"]))
    lat = torch.from_numpy(z["latents"]).cuda()
    outs = []
    for flag in ("2", "0"):
        monkeypatch.setenv("WFD_FUSE_APPLY", flag)
        torch.manual_seed(int(z["seed"]))
        net = AutoencoderTile(**cfg).eval().cuda()
        with torch.no_grad():
            outs.append(net.decode(lat / 0.18215).sample.float().cpu())
        del net
    d = _rel(outs[0], outs[1])
    print(name, "fused vs standalone GroupNorm apply: logits rel-L2", d, "bitwise equal:", bool(torch.equal(outs[0], outs[1])))
    assert d < 1e-3


@pytest.mark.parametrize("name", ["tiny_step.npz", "ice_c1_step.npz"])
def test_softmax_backward_in_the_epilogue_matches_the_two_kernel_form(name, monkeypatch):
    """Attention forward (attention.py:354-368): the fp32 row softmax runs inside two passes of the score product (row
    statistics, then probabilities) instead of an fp32 N x N score matrix and a separate kernel.
    Attention backward (attention.py:367 differentiated): dS = P * (dP - rowdot) / sqrt(C) is written by the epilogue of
    dP = dO V^T with rowdot = sum_c dO * O, instead of an fp32 dP round trip and a separate kernel whose rowdot is
    sum_j dP * P. Same quantity, different summation: the latent gradients agree far below the parity tolerance."""
    z = np.load(os.path.join(GOLD, name))
    cfg = json.loads(str(z["cfg_json"]))
    if name.startswith("ice"):
        mx, my = load_wfc_mats("ice_64x64")
    else:
        T = int(z["T"])
        mats = np.unpackbits(z["mats"])[: 2 * T * T].reshape(2, T, T).astype(bool)
        mx, my = mats[0], mats[1]
    grads = []
    for flag in ("0", "1"):
        monkeypatch.setenv("WFD_NO_FUSE_ATTN_BWD", flag)
        monkeypatch.setenv("WFD_NO_FUSE_ATTN_FWD", flag)
        torch.manual_seed(int(z["seed"]))
        net = AutoencoderTile(**cfg).eval().cuda()
        wfc_loss = WFCLossEinsum(mx, my).cuda()
        lat = torch.from_numpy(z["latents"]).cuda().requires_grad_()
        loss = GuidanceLossFunction.apply(lat, net._native, wfc_loss.native_handle(cfg["out_channels"], lat.device),
                                          1.0 / 0.18215, 5.0)
        (g,) = torch.autograd.grad(loss.sum(), lat)
        grads.append(g.cpu())
        print(name, "WFD_NO_FUSE_ATTN=%s: latent gradient vs the reference's fp64 rel-L2" % flag, _rel(g.cpu(), z["grad_f64"]))
        assert _rel(g.cpu(), z["grad_f64"]) < TOL_GRAD
        del net, wfc_loss
    d = _rel(grads[0], grads[1])
    print(name, "fused vs two-kernel softmax backward: latent gradient rel-L2", d)
    # Two 16-bit computations of the same gradient sit at the same distance from the fp64 reference (measured on B200:
    # 1.18e-2 fused, 1.16e-2 two-kernel at ice config 1) and, their rounding noise being uncorrelated, about sqrt(2) times
    # that from each other: the fused form must not be further from the reference than the two-kernel form by more than
    # a quarter, and the two must stay within twice the reference distance of each other.
    e_fused, e_two = _rel(grads[0], z["grad_f64"]), _rel(grads[1], z["grad_f64"])
    assert e_fused < 1.25 * e_two + 1e-3
    assert d < 2.0 * max(e_fused, e_two)


def test_ice_config1_against_reference_golden():
    """BASELINE config 1: ice 64x64, batch 1, tilenet_sc_space_64x64, seed-0 weights, strength 5."""
    z = np.load(os.path.join(GOLD, "ice_c1_step.npz"))
    cfg = json.loads(str(z["cfg_json"]))
    mx, my = load_wfc_mats("ice_64x64")
    torch.manual_seed(int(z["seed"]))
    net = AutoencoderTile(**cfg).eval().cuda()
    wfc_loss = WFCLossEinsum(mx, my).cuda()
    lat = torch.from_numpy(z["latents"]).cuda().requires_grad_()
    handle = wfc_loss.native_handle(cfg["out_channels"], lat.device)
    loss = GuidanceLossFunction.apply(lat, net._native, handle, 1.0 / 0.18215, 5.0)
    (grad,) = torch.autograd.grad(loss.sum(), lat)
    print("ice C1: loss", loss.item(), "ref", z["loss_f64"], "grad rel-L2", _rel(grad.cpu(), z["grad_f64"]))
    assert _rel(loss.cpu(), z["loss_f64"]) < TOL_LOSS
    assert _rel(grad.cpu(), z["grad_f64"]) < TOL_GRAD
    # logits at the sampled pixels
    plan = net._native.plan(1, 64, 64, lat.device)
    logits = plan.logits_view(1).float()  # (1, T_pad, H, W) view
    got = logits[0].reshape(cfg["out_channels"], -1)[:, torch.from_numpy(z["pix"]).cuda()].t().cpu()
    assert _rel(got, z["logits_pix_f64"]) < TOL_LOGITS
    # argmax: bit-exact on the identical wave; against the reference's own tiles only cells with a clear margin count
    wave = handle.wave_view(1, 64, 64)
    tiles = argmax_tiles(wave)
    assert np.array_equal(tiles.cpu().numpy(), np.argmax(wave.float().cpu().numpy(), axis=-1))
    clear = z["top2_margin_f64"].reshape(64, 64) > 0.05
    assert (tiles[0].cpu().numpy()[clear] == z["tiles_f64"][clear]).mean() > 0.99


def test_ice_config2_batch8_against_reference_golden():
    """BASELINE config 2's benchmarked shape: ice 64x64, BATCH 8, against the unmodified reference's fp64 loss (8,) and
    latent gradient; slot 0 carries the config-1 latents and must reproduce the batch-1 run of this plan bit for bit."""
    z = np.load(os.path.join(GOLD, "ice_c2_step.npz"))
    cfg = json.loads(str(z["cfg_json"]))
    mx, my = load_wfc_mats("ice_64x64")
    torch.manual_seed(int(z["seed"]))
    net = AutoencoderTile(**cfg).eval().cuda()
    wfc_loss = WFCLossEinsum(mx, my).cuda()
    handle = wfc_loss.native_handle(cfg["out_channels"], torch.device("cuda:0"))
    lat = torch.from_numpy(z["latents"]).cuda()
    x = lat.clone().requires_grad_()
    loss = GuidanceLossFunction.apply(x, net._native, handle, 1.0 / 0.18215, 5.0)
    (grad,) = torch.autograd.grad(loss.sum(), x)
    rel_l = np.abs(loss.detach().cpu().numpy().astype(np.float64) - z["loss_f64"]) / np.abs(z["loss_f64"])
    rel_g = [_rel(grad[b].cpu(), z["grad_f64"][b]) for b in range(8)]
    print("ice C2: loss rel max", rel_l.max(), "grad rel-L2 per map", ["%.3e" % r for r in rel_g])
    assert rel_l.max() < TOL_LOSS
    assert max(rel_g) < TOL_GRAD and _rel(grad.cpu(), z["grad_f64"]) < TOL_GRAD
    assert loss[5].item() == loss[2].item() and torch.equal(grad[5], grad[2])
    # the same plan on the config-1 latents alone: bitwise the batch's slot 0 (integer cross-CTA sums, fixed tile walk)
    x1 = lat[:1].clone().requires_grad_()
    loss1 = GuidanceLossFunction.apply(x1, net._native, handle, 1.0 / 0.18215, 5.0)
    (grad1,) = torch.autograd.grad(loss1.sum(), x1)
    assert loss1.item() == loss[0].item() and torch.equal(grad1[0], grad[0])
    c1 = np.load(os.path.join(GOLD, "ice_c1_step.npz"))
    assert abs(loss1.item() - float(np.asarray(c1["loss_f64"]).reshape(-1)[0])) < TOL_LOSS * loss1.item()


def test_fp16_operand_mode_small(monkeypatch):
    """WFD_DTYPE16=fp16 (the reference's set_precision('half') operand type) through the whole step: same tolerances
    for logits and loss; the gradient is held to the same bound thanks to the power-of-two gradient scale of fp16 plans."""
    monkeypatch.setenv("WFD_DTYPE16", "fp16")
    z, cfg, mats, net = _load_small("tiny_step.npz")
    wfc_loss = WFCLossEinsum(mats[0], mats[1]).cuda()
    lat = torch.from_numpy(z["latents"]).cuda().requires_grad_()
    loss = GuidanceLossFunction.apply(lat, net._native, wfc_loss.native_handle(cfg["out_channels"], lat.device),
                                      1.0 / 0.18215, 5.0)
    (grad,) = torch.autograd.grad(loss.sum(), lat)
    with torch.no_grad():
        logits = net.decode(lat.detach() / 0.18215).sample
    assert _rel(logits.cpu(), z["logits_f64"]) < TOL_LOGITS
    assert _rel(loss.cpu(), z["loss_f64"]) < TOL_LOSS
    assert _rel(grad.cpu(), z["grad_f64"]) < TOL_GRAD


ALL_TILESETS = ["ashworld_64x64", "badlands_64x64", "desert_64x64", "ice_64x64", "install_64x64", "jungle_64x64",
                "platform_32x32", "platform_64x64", "twilight_64x64"]


@pytest.mark.parametrize("tileset", ALL_TILESETS)
def test_wfc_known_answers_on_device(tileset):
    mx, my = load_wfc_mats(tileset)
    T = mx.shape[0]
    t_pad = round_up_channels(T)
    wfc_loss = WFCLossEinsum(mx, my).cuda()
    H, W = 16, 8
    # uniform wave over t_pad channels: analytic value
    wave = torch.full((2, t_pad, H, W), 1.0 / t_pad, device="cuda")
    got = wfc_loss(wave).cpu().numpy()
    # the kernels see 1/t_pad rounded to bf16; the loss is quadratic in the wave, so the analytic value scales exactly
    r = (torch.tensor(1.0 / t_pad).bfloat16().float().item() * t_pad) ** 2
    assert np.allclose(got, orc.uniform_wave_loss(mx, my, t_pad) * r, rtol=2e-4)
    # zero logits through the fused softmax give the same rounded wave
    got = wfc_loss(torch.zeros(2, t_pad, H, W, device="cuda"), softmax=True).cpu().numpy()
    assert np.allclose(got, orc.uniform_wave_loss(mx, my, t_pad) * r, rtol=2e-4)
    # one-hot wave of an integer tile map: exactly the fraction of forbidden adjacent pairs
    rng = np.random.RandomState(0)
    tiles = rng.randint(0, T, size=(H, W))
    onehot = torch.zeros(1, t_pad, H, W)
    onehot[0, torch.from_numpy(tiles), torch.arange(H)[:, None], torch.arange(W)[None, :]] = 1.0
    got = wfc_loss(onehot.cuda()).item()
    assert abs(got - orc.forbidden_pair_fraction(tiles, mx, my)) < 1e-6
    # a map whose pairs are all allowed has loss exactly 0: walk allowed transitions along a single row
    ax = mx.copy()
    row = [int(np.argwhere(ax.any(1))[0, 0])]
    for _ in range(W - 1):
        nxt = np.argwhere(ax[row[-1]])
        if len(nxt) == 0:
            break
        row.append(int(nxt[0, 0]))
    if len(row) == W:
        oh = torch.zeros(1, t_pad, 1, W)
        oh[0, torch.tensor(row), 0, torch.arange(W)] = 1.0
        # a 1 x W map cannot be tiled into 128-cell patches: replicate the row only if the vertical pairs are allowed
        col_ok = all(my[t, t] for t in row)
        if col_ok:
            oh = oh.expand(1, t_pad, 16, W).contiguous()
            assert wfc_loss(oh.cuda()).item() == 0.0


def test_wfc_loss_accepts_any_channel_count_like_the_reference():
    """losses.py:63-66 slices the first `count` channels of whatever it is given: a wave with exactly `count` channels (not
    a multiple of 64) must work, forward and backward, with and without the softmax."""
    mx, my = load_wfc_mats("install_64x64")
    T = mx.shape[0]
    assert T % 64 != 0
    wfc_loss = WFCLossEinsum(mx, my).cuda()
    g = torch.Generator().manual_seed(5)
    logits = torch.randn(2, T, 16, 8, generator=g)
    wave = logits.softmax(dim=1)
    want = orc.wfc_loss(wave.double(), mx, my)
    w = wave.cuda().requires_grad_()
    got = wfc_loss(w)
    assert got.shape == (2,) and _rel(got.cpu(), want) < TOL_LOSS
    (gw,) = torch.autograd.grad(got.sum(), w)
    assert gw.shape == w.shape
    wd = wave.double().requires_grad_()
    (gref,) = torch.autograd.grad(orc.wfc_loss(wd, mx, my).sum(), wd)
    assert _rel(gw.cpu(), gref) < TOL_GRAD
    got_sm = wfc_loss(logits.cuda(), softmax=True)
    assert _rel(got_sm.cpu(), want) < TOL_LOSS


def test_wfc_gradient_matches_closed_form():
    mx, my = load_wfc_mats("install_64x64")
    T = mx.shape[0]
    t_pad = round_up_channels(T)
    wfc_loss = WFCLossEinsum(mx, my).cuda()
    torch.manual_seed(0)
    wave = torch.softmax(torch.randn(2, t_pad, 8, 16) * 3, dim=1)
    wave16 = wave.bfloat16().float()
    w = wave16.cuda().requires_grad_()
    loss = wfc_loss(w)
    (g,) = torch.autograd.grad((loss * torch.tensor([1.0, 2.0], device="cuda")).sum(), w)
    want_loss = orc.wfc_loss(wave16.double(), mx, my)
    want_g = orc.wfc_grad_closed_form(wave16.double(), mx, my) * torch.tensor([1.0, 2.0]).view(2, 1, 1, 1)
    assert _rel(loss.detach().cpu(), want_loss) < 1e-4
    assert _rel(g.cpu(), want_g) < 5e-3
    assert g.shape == w.shape and g[:, T:].abs().max().item() == 0.0


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_argmax_numpy_semantics(dtype):
    torch.manual_seed(0)
    x = torch.randn(37, 3200).to(dtype)
    x[3, 100] = x[3, 2000] = 50.0           # tie: first index wins
    x[4, 3199] = 60.0                        # last element
    x[5, 0] = 60.0                           # first element
    x[6, :] = 1.0                            # all equal -> 0
    if dtype != torch.float16:
        x[7, 77] = float("nan")              # NaN is maximal, first NaN wins
        x[7, 99] = float("nan")
        x[8, 5] = float("inf")
    got = argmax_tiles(x.cuda()).cpu().numpy()
    want = np.argmax(x.float().numpy(), axis=-1)
    assert got.dtype == np.int64 and np.array_equal(got, want)
    # ragged row length (not a multiple of the 32-lane vector stride) and an empty batch
    y = torch.randn(5, 7, 1032).to(dtype)
    assert np.array_equal(argmax_tiles(y.cuda()).cpu().numpy(), np.argmax(y.float().numpy(), axis=-1))
    assert argmax_tiles(torch.zeros(0, 128, dtype=dtype, device="cuda")).shape == (0,)
    # rows that do not start 16-byte aligned (the ice tileset's 3074 real tiles cut out of the 3200 padded channels)
    w = torch.randn(9, 3074).to(dtype)
    assert np.array_equal(argmax_tiles(w.cuda()).cpu().numpy(), np.argmax(w.float().numpy(), axis=-1))


def test_batch_properties_full_size():
    """Size-independent properties at BASELINE config 2's shape (ice 64x64, batch 8): maps are independent, the loss is
    linear in the strength, and a batch equals its single-map runs."""
    mx, my = load_wfc_mats("ice_64x64")
    t_pad = round_up_channels(mx.shape[0])
    from wavefunctiondiffusion_b200.utils.tilesets import tilenet_config

    torch.manual_seed(0)
    net = AutoencoderTile(**tilenet_config("64x64", t_pad)).eval().cuda()
    wfc_loss = WFCLossEinsum(mx, my).cuda()
    handle = wfc_loss.native_handle(t_pad, torch.device("cuda:0"))
    g = torch.Generator().manual_seed(5)
    lat = torch.randn(8, 4, 64, 64, generator=g).cuda()
    lat[5] = lat[2]
    z = lat.clone().requires_grad_()
    loss = GuidanceLossFunction.apply(z, net._native, handle, 1 / 0.18215, 5.0)
    (grad,) = torch.autograd.grad(loss.sum(), z)
    assert torch.isfinite(loss).all() and torch.isfinite(grad).all()
    # cross-CTA reductions are fixed-point integer atomics (order independent): identical maps give bitwise identical
    # results whatever their batch slot, and a second run reproduces the first bit for bit
    assert loss[5].item() == loss[2].item()
    assert torch.equal(grad[5], grad[2])
    z2 = lat.clone().requires_grad_()
    loss_b = GuidanceLossFunction.apply(z2, net._native, handle, 1 / 0.18215, 5.0)
    (grad_b,) = torch.autograd.grad(loss_b.sum(), z2)
    assert torch.equal(loss_b, loss) and torch.equal(grad_b, grad)
    z1 = lat[:1].clone().requires_grad_()
    loss1 = GuidanceLossFunction.apply(z1, net._native, handle, 1 / 0.18215, 10.0)
    (grad1,) = torch.autograd.grad(loss1.sum(), z1)
    assert abs(loss1.item() - 2 * loss[0].item()) <= 1e-4 * abs(loss1.item())
    assert _rel(grad1[0].cpu(), 2 * grad[0].cpu()) < TOL_GRAD


def test_rule_violation_count_of_tile_maps():
    """Integer work, bit-exact: forbidden adjacent pairs of argmax tile maps against the oracle's count (which is the
    reference loss of the one-hot wave times the pair count), including ids in the padded channels and ragged sizes."""
    rng = np.random.RandomState(9)
    T = 100
    mats = rng.rand(2, T, T) < 0.3
    loss = WFCLossEinsum(mats[0], mats[1]).cuda()
    for (B, H, W) in ((1, 1, 2), (2, 7, 5), (3, 64, 64)):
        tiles = rng.randint(0, T, size=(B, H, W)).astype(np.int64)
        got = loss.count_violations(torch.from_numpy(tiles).cuda()).cpu().numpy()
        for b in range(B):
            P = H * (W - 1) + (H - 1) * W
            want = orc.forbidden_pair_fraction(tiles[b], mats[0], mats[1]) * P
            assert got[b].sum() == round(want)
            assert got[b, 0] == np.count_nonzero(~mats[0][tiles[b][:, :-1], tiles[b][:, 1:]])
            assert got[b, 1] == np.count_nonzero(~mats[1][tiles[b][:-1], tiles[b][1:]])
    # ids in the padded channels (>= T) can come out of an argmax over T_pad channels: every pair they touch violates
    tiles = np.zeros((1, 2, 3), dtype=np.int64)
    tiles[0, 0, 1] = T + 5
    allow = np.ones((2, T, T), dtype=bool)
    got = WFCLossEinsum(allow[0], allow[1]).cuda().count_violations(torch.from_numpy(tiles).cuda()).cpu().numpy()
    assert got.tolist() == [[2, 1]]
    # (H, W) input is treated as a batch of one
    assert loss.count_violations(torch.zeros(4, 4, dtype=torch.int64, device="cuda")).shape == (1, 2)


def test_tile_post_processing_is_bit_exact():
    """Integer work after the argmax (ui.py:228-243): mirror symmetry + gen_to_sc[shrink_to_gen[.]] on the device against
    the oracle's numpy restatement, including ids in the padded channels (counted, mapped to 0xFFFF) and ragged sizes."""
    from wavefunctiondiffusion_b200.scmap_ops import TileTables, finish_tiles, wave_to_map

    rng = np.random.RandomState(3)
    T, T_gen = 307, 449
    s2g = rng.permutation(T_gen)[:T].astype(np.int64)
    g2s = rng.randint(0, 32624, size=T_gen).astype(np.int32)
    mapping = rng.permutation(T).astype(np.int64)
    tables = TileTables(s2g, g2s, mapping, "cuda")
    for (B, H, W) in ((1, 1, 1), (2, 7, 5), (3, 64, 64)):
        tiles = rng.randint(0, T, size=(B, H, W)).astype(np.int64)
        t_out, sc, bad = finish_tiles(torch.from_numpy(tiles).cuda(), tables, symmetry=False)
        assert bad.item() == 0 and torch.equal(t_out.cpu(), torch.from_numpy(tiles))
        assert np.array_equal(sc.cpu().numpy().astype(np.int64), np.stack([orc.tiles_to_sc(t, s2g, g2s) for t in tiles]))
        t_out, sc, bad = finish_tiles(torch.from_numpy(tiles).cuda(), tables, symmetry=True)
        want_t = np.stack([orc.add_symmetry(t, mapping) for t in tiles])
        assert bad.item() == 0 and np.array_equal(t_out.cpu().numpy(), want_t) and t_out.shape == (B, H, 2 * W)
        assert np.array_equal(sc.cpu().numpy().astype(np.int64), np.stack([orc.tiles_to_sc(t, s2g, g2s) for t in want_t]))
    # the shipped ice tables and outputs of the reference's own add_symmetry / id chain (oracle/make_golden.py)
    z = np.load(os.path.join(GOLD, "tiles_post.npz"))
    ice = TileTables(z["shrink_to_gen"], z["gen_to_sc"], z["mapping"], "cuda")
    t_out, sc, bad = finish_tiles(torch.from_numpy(z["tiles"].astype(np.int64)).cuda(), ice, symmetry=True)
    assert bad.item() == 0 and np.array_equal(t_out[0].cpu().numpy(), z["mirrored"])
    assert np.array_equal(sc[0].cpu().numpy().astype(np.int64), z["sc"].astype(np.int64))
    # an id from the padded channels: numpy would raise, the kernel flags it
    tiles = np.zeros((1, 2, 3), dtype=np.int64)
    tiles[0, 1, 2] = T + 3
    _, sc, bad = finish_tiles(torch.from_numpy(tiles).cuda(), tables)
    assert bad.item() == 1 and sc[0, 1, 2].item() == 0xFFFF and sc[0, 0, 0].item() == g2s[s2g[0]]
    # fused: wave -> argmax over the first T channels -> map
    wave = torch.rand(2, 8, 8, 384, device="cuda").bfloat16()
    t_out, sc, bad = wave_to_map(wave, tables, count=T)
    want = np.argmax(wave[..., :T].float().cpu().numpy(), axis=-1)
    assert bad.item() == 0 and np.array_equal(t_out.cpu().numpy(), want)
    assert np.array_equal(sc.cpu().numpy().astype(np.int64), np.stack([orc.tiles_to_sc(t, s2g, g2s) for t in want]))
