"""GPU: the sparse form of the loss contraction (csrc/wfc_sparse.cu, SURVEY 8f-4: p^T (1 - A) q = sum(p) sum(q) - p^T A q with A
as lists of allowed pairs) against the same known answers, closed-form gradient and reference goldens as the dense
tensor-core path, and against the dense path itself.

Tolerances as tests/test_parity_gpu.py: loss rel <= 1e-3, latent gradient rel-L2 <= 5e-2 against the reference's fp64;
sparse against dense on the same 16-bit wave: loss rel <= 1e-5, gradient rel-L2 <= 5e-3 (both store a 16-bit `a`)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import wfd_oracle as orc
from wavefunctiondiffusion_b200.utils.tilesets import load_wfc_mats, round_up_channels
from wavefunctiondiffusion_b200.wf_diffusers.native_ops import GuidanceLossFunction
from wavefunctiondiffusion_b200.wf_diffusers.wf_vae import AutoencoderTile
from wavefunctiondiffusion_b200.wfdf.losses import WFCLossEinsum

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ALL_TILESETS = ["ashworld_64x64", "badlands_64x64", "desert_64x64", "ice_64x64", "install_64x64", "jungle_64x64",
                "platform_32x32", "platform_64x64", "twilight_64x64"]


def _rel(a, b):
    a = torch.as_tensor(np.asarray(a.detach() if torch.is_tensor(a) else a), dtype=torch.float64)
    b = torch.as_tensor(np.asarray(b.detach() if torch.is_tensor(b) else b), dtype=torch.float64)
    return ((a - b).norm() / b.norm()).item()


@pytest.mark.parametrize("tileset", ALL_TILESETS)
def test_known_answers(tileset):
    mx, my = load_wfc_mats(tileset)
    T = mx.shape[0]
    t_pad = round_up_channels(T)
    loss = WFCLossEinsum(mx, my).cuda().set_sparse(True)
    H, W = 16, 8
    wave = torch.full((2, t_pad, H, W), 1.0 / t_pad, device="cuda")
    r = (torch.tensor(1.0 / t_pad).bfloat16().float().item() * t_pad) ** 2
    assert np.allclose(loss(wave).cpu().numpy(), orc.uniform_wave_loss(mx, my, t_pad) * r, rtol=2e-4)
    assert np.allclose(loss(torch.zeros(2, t_pad, H, W, device="cuda"), softmax=True).cpu().numpy(),
                       orc.uniform_wave_loss(mx, my, t_pad) * r, rtol=2e-4)
    rng = np.random.RandomState(0)
    tiles = rng.randint(0, T, size=(H, W))
    onehot = torch.zeros(1, t_pad, H, W)
    onehot[0, torch.from_numpy(tiles), torch.arange(H)[:, None], torch.arange(W)[None, :]] = 1.0
    assert abs(loss(onehot.cuda()).item() - orc.forbidden_pair_fraction(tiles, mx, my)) < 1e-6


@pytest.mark.parametrize("shape", [(8, 16), (7, 5), (1, 9), (3, 1), (2, 2)])
def test_gradient_matches_closed_form_any_map_shape(shape):
    """Strips of four cells with zero rows outside the map: widths that are no multiple of four, single rows / columns
    (the dense path needs 128-cell patches; the reference takes any shape, losses.py:63-75)."""
    H, W = shape
    mx, my = load_wfc_mats("install_64x64")
    T = mx.shape[0]
    t_pad = round_up_channels(T)
    loss_mod = WFCLossEinsum(mx, my).cuda().set_sparse(True)
    torch.manual_seed(H * 31 + W)
    wave16 = torch.softmax(torch.randn(2, t_pad, H, W) * 3, dim=1).bfloat16().float()
    w = wave16.cuda().requires_grad_()
    loss = loss_mod(w)
    (g,) = torch.autograd.grad((loss * torch.tensor([1.0, 2.0], device="cuda")).sum(), w)
    want_loss = orc.wfc_loss(wave16.double(), mx, my)
    want_g = orc.wfc_grad_closed_form(wave16.double(), mx, my) * torch.tensor([1.0, 2.0]).view(2, 1, 1, 1)
    assert _rel(loss.detach().cpu(), want_loss) < 1e-4
    assert _rel(g.cpu(), want_g) < 5e-3
    assert g[:, T:].abs().max().item() == 0.0


def test_sparse_against_dense_full_size_ice():
    """ice 64 x 64, batch 2, softmax inside: both contractions see the same 16-bit wave."""
    mx, my = load_wfc_mats("ice_64x64")
    t_pad = round_up_channels(mx.shape[0])
    g = torch.Generator(device="cuda").manual_seed(2)
    logits = (3.0 * torch.randn(2, t_pad, 64, 64, generator=g, device="cuda")).requires_grad_()
    out = {}
    for name in ("dense", "sparse"):
        mod = WFCLossEinsum(mx, my).cuda().set_sparse(name == "sparse")
        loss = mod(logits, softmax=True)
        (gr,) = torch.autograd.grad(loss.sum(), logits)
        out[name] = (loss.detach().double().cpu(), gr.double().cpu())
    print("sparse vs dense: loss", out["sparse"][0].tolist(), out["dense"][0].tolist(), "grad rel", _rel(out["sparse"][1], out["dense"][1]))
    assert _rel(out["sparse"][0], out["dense"][0]) < 1e-5
    assert _rel(out["sparse"][1], out["dense"][1]) < 5e-3


def test_ice_config1_golden_through_the_sparse_contraction(monkeypatch):
    """BASELINE config 1 against the unmodified reference's fp64 loss and latent gradient, the fused guidance step scoring
    through the sparse contraction (WFD_WFC_SPARSE=1 switches every new handle)."""
    monkeypatch.setenv("WFD_WFC_SPARSE", "1")
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
    print("ice C1 sparse: loss", loss.item(), "ref", z["loss_f64"], "grad rel-L2", _rel(grad.cpu(), z["grad_f64"]))
    assert _rel(loss.cpu(), z["loss_f64"]) < 1e-3
    assert _rel(grad.cpu(), z["grad_f64"]) < 5e-2
