"""CPU: the oracle restatement against the golden vectors produced by the unmodified reference (oracle/make_golden.py)
and against analytic known-answer values of the loss."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import wfd_oracle as orc
from wavefunctiondiffusion_b200.utils.tilesets import load_wfc_mats, round_up_channels, tilenet_config
from wavefunctiondiffusion_b200.wf_diffusers.wf_vae import AutoencoderTile

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load_small(name):
    z = np.load(os.path.join(GOLD, name))
    cfg = json.loads(str(z["cfg_json"]))
    T = int(z["T"])
    mats = np.unpackbits(z["mats"])[: 2 * T * T].reshape(2, T, T).astype(bool)
    torch.manual_seed(int(z["seed"]))
    net = AutoencoderTile(**cfg)
    return z, cfg, mats, net


@pytest.mark.parametrize("name", ["tiny_step.npz", "tiny_down_step.npz"])
def test_oracle_matches_reference_small(name):
    z, cfg, mats, net = _load_small(name)
    lat = torch.from_numpy(z["latents"])
    r = orc.guidance_step(net.state_dict(), cfg, mats[0], mats[1], lat, float(z["strength"]), torch.float64)
    assert np.allclose(r["logits"].numpy(), z["logits_f64"], rtol=0, atol=1e-10)
    assert np.allclose(r["loss"].numpy(), z["loss_f64"], rtol=1e-12)
    g = z["grad_f64"]
    assert np.linalg.norm(r["grad"].numpy() - g) <= 1e-10 * np.linalg.norm(g)
    r32 = orc.guidance_step(net.state_dict(), cfg, mats[0], mats[1], lat, float(z["strength"]), torch.float32)
    assert np.allclose(r32["loss"].numpy(), z["loss_f64"], rtol=2e-6)
    assert np.linalg.norm(r32["grad"].numpy() - g) <= 2e-4 * np.linalg.norm(g)


def test_loss_three_ways_and_closed_form_gradient():
    rng = np.random.RandomState(0)
    T, Tp = 37, 48
    mats = rng.rand(2, T, T) < 0.1
    p = torch.softmax(torch.randn(2, Tp, 5, 7, dtype=torch.float64), dim=1).requires_grad_()
    a = orc.wfc_loss(p, mats[0], mats[1])
    b = orc.wfc_loss_sparse(p.detach(), mats[0], mats[1])
    assert torch.allclose(a.detach(), b, rtol=1e-12)
    (g,) = torch.autograd.grad(a.sum() * 3.0, p)
    g2 = orc.wfc_grad_closed_form(p.detach(), mats[0], mats[1], 3.0)
    assert torch.allclose(g, g2, rtol=1e-10, atol=1e-14)


@pytest.mark.parametrize("tileset", ["ice_64x64", "install_64x64", "platform_32x32"])
def test_uniform_and_one_hot_known_answers(tileset):
    mx, my = load_wfc_mats(tileset)
    T = mx.shape[0]
    t_pad = round_up_channels(T)
    from wavefunctiondiffusion_b200.utils.tilesets import WFC_DATA_DIR

    kat = float(np.load(os.path.join(WFC_DATA_DIR, tileset + ".npz"))["uniform_kat"])
    assert abs(orc.uniform_wave_loss(mx, my, t_pad) - kat) < 1e-12
    H = W = 6
    wave = torch.full((1, t_pad, H, W), 1.0 / t_pad, dtype=torch.float64)
    assert abs(orc.wfc_loss_sparse(wave, mx, my).item() - kat) < 1e-10
    rng = np.random.RandomState(1)
    tiles = rng.randint(0, T, size=(H, W))
    onehot = torch.zeros(1, t_pad, H, W, dtype=torch.float64)
    onehot[0, torch.from_numpy(tiles), torch.arange(H)[:, None], torch.arange(W)[None, :]] = 1.0
    want = orc.forbidden_pair_fraction(tiles, mx, my)
    assert abs(orc.wfc_loss_sparse(onehot, mx, my).item() - want) < 1e-12


def test_ice_ids_documented_in_survey():
    mx, my = load_wfc_mats("ice_64x64")
    assert mx.shape == (3074, 3074) and mx.sum() == 19182 and my.sum() == 25476
    assert not mx[0].any() and not mx[:, 0].any()


@pytest.mark.slow
def test_oracle_matches_reference_ice_c1():
    """BASELINE config 1 (ice 64x64, batch 1): ~10 s of fp32 CPU work."""
    z = np.load(os.path.join(GOLD, "ice_c1_step.npz"))
    cfg = json.loads(str(z["cfg_json"]))
    mx, my = load_wfc_mats("ice_64x64")
    torch.manual_seed(int(z["seed"]))
    net = AutoencoderTile(**cfg)
    r = orc.guidance_step(net.state_dict(), cfg, mx, my, torch.from_numpy(z["latents"]), 5.0, torch.float32)
    assert np.allclose(r["loss"].numpy(), z["loss_f64"], rtol=2e-6)
    g = z["grad_f64"]
    assert np.linalg.norm(r["grad"].numpy() - g) <= 1e-3 * np.linalg.norm(g)
    tiles = orc.argmax_tiles(r["wave"].permute(0, 2, 3, 1).numpy())[0]
    assert (tiles != z["tiles_f64"]).mean() < 0.002


@pytest.mark.slow
def test_oracle_matches_reference_ice_c2_batch8_slot():
    """BASELINE config 2's shape (batch 8): the golden run of the unmodified reference on the whole batch against the
    oracle on two of its maps (maps are independent: the loss is per map, GroupNorm and attention are per sample)."""
    z = np.load(os.path.join(GOLD, "ice_c2_step.npz"))
    c1 = np.load(os.path.join(GOLD, "ice_c1_step.npz"))
    cfg = json.loads(str(z["cfg_json"]))
    mx, my = load_wfc_mats("ice_64x64")
    assert z["latents"].shape == (8, 4, 64, 64) and np.array_equal(z["latents"][0], c1["latents"][0])
    assert np.array_equal(z["latents"][5], z["latents"][2])
    assert abs(z["loss_f64"][0] - float(np.asarray(c1["loss_f64"]).reshape(-1)[0])) < 1e-9
    assert z["loss_f64"][5] == z["loss_f64"][2]
    torch.manual_seed(int(z["seed"]))
    net = AutoencoderTile(**cfg)
    lat = torch.from_numpy(z["latents"][[3, 7]])
    r = orc.guidance_step(net.state_dict(), cfg, mx, my, lat, 5.0, torch.float32)
    assert np.allclose(r["loss"].numpy(), z["loss_f64"][[3, 7]], rtol=2e-6)
    g = z["grad_f64"][[3, 7]].astype(np.float64)
    assert np.linalg.norm(r["grad"].numpy() - g) <= 1e-3 * np.linalg.norm(g)


def test_oracle_encode_matches_reference():
    """AutoencoderTile.encode (the img2img wave input): oracle restatement against the reference's moments."""
    z = np.load(os.path.join(GOLD, "tiny_encode.npz"))
    cfg = json.loads(str(z["cfg_json"]))
    torch.manual_seed(int(z["seed"]))
    net = AutoencoderTile(**cfg)
    onehot = torch.nn.functional.one_hot(torch.from_numpy(z["tiles"]), num_classes=cfg["in_channels"]).permute(0, 3, 1, 2)
    m = orc.encode(net.state_dict(), cfg, onehot, torch.float64)
    assert m.shape == (2, 8, 16, 16) and np.allclose(m.numpy(), z["moments_onehot_f64"], rtol=0, atol=1e-10)
    m = orc.encode(net.state_dict(), cfg, torch.from_numpy(z["soft"]), torch.float64)
    assert np.allclose(m.numpy(), z["moments_soft_f64"], rtol=0, atol=1e-10)
    # the product's eager encoder modules (CPU tensors) are the same arithmetic
    with torch.no_grad():
        eager = net.encode(onehot.float()).latent_dist.parameters
    assert np.allclose(eager.numpy(), z["moments_onehot_f32"], atol=1e-5)


def test_tile_post_processing_against_the_reference():
    """add_symmetry and the id chain of the oracle against outputs of the reference's own functions (tiles_post.npz)."""
    z = np.load(os.path.join(GOLD, "tiles_post.npz"))
    mirrored = orc.add_symmetry(z["tiles"], z["mapping"])
    assert mirrored.shape == (12, 18) and np.array_equal(mirrored, z["mirrored"])
    assert np.array_equal(orc.tiles_to_sc(mirrored, z["shrink_to_gen"], z["gen_to_sc"]), z["sc"])


def test_ddim_coefficients_shape():
    a, b = orc.ddim_coefficients(50)
    assert a.shape == (50,) and b.shape == (50,) and (a > 1.0).all() and a[0] > a[-1] and (b < 0).all()
