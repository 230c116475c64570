"""GPU: the pipeline mirror (API surface, guidance block with the scheduler's chain factor, precision switch, micro-batching)
and the larger map sizes."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import wfd_oracle as orc
from wavefunctiondiffusion_b200.utils.harness import (AffineDDIMScheduler, ConstantTextEncoder, RandomNoiseUNet, TinyImageVAE,
                                                      WhitespaceTokenizer)
from wavefunctiondiffusion_b200.utils.tilesets import load_wfc_mats, round_up_channels, tilenet_config
from wavefunctiondiffusion_b200.wf_diffusers.native_ops import GuidanceLossFunction
from wavefunctiondiffusion_b200.wf_diffusers.pipeline_wfd import WaveFunctionDiffusionPipeline, WaveFunctionDiffusionPipelineOutput
from wavefunctiondiffusion_b200.wf_diffusers.wf_vae import AutoencoderTile
from wavefunctiondiffusion_b200.wfdf.losses import WFCLossEinsum

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return ((a - b).norm() / b.norm()).item()


def _tiny_pipeline(tmp_path, latent=16):
    z = np.load(os.path.join(GOLD, "tiny_step.npz"))
    cfg = json.loads(str(z["cfg_json"]))
    T = int(z["T"])
    mats = np.unpackbits(z["mats"])[: 2 * T * T].reshape(2, T, T).astype(bool)
    torch.manual_seed(int(z["seed"]))
    net = AutoencoderTile(**cfg).eval()
    npz = str(tmp_path / "wfc.npz")
    np.savez(npz, wfc_mats=mats)
    pipe = WaveFunctionDiffusionPipeline(TinyImageVAE(), ConstantTextEncoder(), WhitespaceTokenizer(), net,
                                         RandomNoiseUNet(4, latent), AffineDDIMScheduler(), None, None, wfc_data_path=npz)
    pipe.set_progress_bar_config(disable=True)
    return pipe.to("cuda"), z, cfg, mats, net


def test_guidance_block_applies_the_scheduler_chain_factor(tmp_path):
    pipe, z, cfg, mats, net = _tiny_pipeline(tmp_path)
    pipe.scheduler.set_timesteps(50, device="cuda")
    t = pipe.scheduler.timesteps[25]
    a, b = pipe.scheduler.coefficients(t)
    lat = torch.from_numpy(z["latents"]).cuda()
    eps = torch.randn(lat.shape, generator=torch.Generator().manual_seed(3)).cuda()
    new = pipe.wfc_guidance(lat, eps, t, 5.0)
    # oracle: loss on lat_prev = a*lat + b*eps; d/d lat = a * d/d lat_prev
    want = orc.guidance_step(net.state_dict(), cfg, mats[0], mats[1], (a * lat + b * eps).cpu(), 5.0, torch.float64)
    assert _rel((lat - new).cpu(), a * want["grad"]) < 5e-2


def test_final_stage_loop_as_one_captured_sequence(tmp_path, monkeypatch):
    """pipeline_wfd.py:626-639 (20 guidance-only iterations on the last noise prediction): the library's fused loop
    (scheduler step, decode, loss, backward, latent update of every iteration in one captured launch sequence) against
    the same loop written as the reference writes it (scheduler.step under autograd, one wfc_guidance call per
    iteration)."""
    pipe, z, cfg, mats, net = _tiny_pipeline(tmp_path)
    pipe.scheduler.set_timesteps(50, device="cuda")
    t = pipe.scheduler.timesteps[-1]
    lat = torch.from_numpy(z["latents"]).cuda()
    eps = torch.randn(lat.shape, generator=torch.Generator().manual_seed(5)).cuda()
    fused = pipe.wfc_guidance_final(lat, eps, t, 10.0, 6)
    monkeypatch.setenv("WFD_FUSED_FINAL", "0")
    plain = pipe.wfc_guidance_final(lat, eps, t, 10.0, 6)
    assert torch.isfinite(fused).all() and (fused - lat).abs().max() > 0
    assert _rel((fused - lat).cpu(), (plain - lat).cpu()) < 2e-2
    # one iteration of the loop is the guidance block itself
    monkeypatch.setenv("WFD_FUSED_FINAL", "1")
    one = pipe.wfc_guidance_final(lat, eps, t, 10.0, 1)
    assert _rel((one - lat).cpu(), (pipe.wfc_guidance(lat, eps, t, 10.0) - lat).cpu()) < 2e-2


def test_call_signature_outputs_and_errors(tmp_path):
    pipe, z, cfg, mats, net = _tiny_pipeline(tmp_path)
    out = pipe(["a b", "c"], height=128, width=128, num_inference_steps=6, guidance_scale=3.5,
               wfc_guidance_start_step=2, wfc_guidance_strength=5, wfc_guidance_final_steps=2,
               wfc_guidance_final_strength=10, output_type="np", generator=torch.Generator("cuda").manual_seed(0))
    assert isinstance(out, WaveFunctionDiffusionPipelineOutput)
    assert out.waves.shape == (2, 16, 16, cfg["out_channels"]) and out.waves.dtype == np.float32
    assert np.allclose(out.waves.sum(-1), 1.0, atol=2e-2)
    assert out.images.shape == (2, 128, 128, 3) and out.nsfw_content_detected is None
    tiles = out.waves[0].argmax(axis=2)
    assert tiles.shape == (16, 16)
    with pytest.raises(ValueError):
        pipe("x", height=100, width=128)
    with pytest.raises(NotImplementedError):
        pipe.set_precision("double")
    pipe.enable_vae_slicing()
    assert pipe.tile_vae.use_slicing
    pipe.disable_vae_slicing()


def test_set_precision_half_keeps_working(tmp_path):
    pipe, z, cfg, mats, net = _tiny_pipeline(tmp_path)
    pipe.set_precision("half")
    assert pipe.tile_vae.dtype == torch.float16 and pipe.wfc_loss.cx.dtype == torch.float16
    lat = torch.from_numpy(z["latents"]).cuda().half()
    wave = pipe.decode_latents_tile(lat)
    assert wave.dtype == torch.float16 and wave.shape == (2, cfg["out_channels"], 16, 16)
    loss = pipe.guidance_loss(lat, 5.0)
    assert _rel(loss.float().cpu(), z["loss_f64"]) < 5e-3     # fp16 latents + half-precision weights
    pipe.set_precision("float")
    assert pipe.tile_vae.dtype == torch.float32


def test_micro_batching_matches_one_call(tmp_path, monkeypatch):
    pipe, z, cfg, mats, net = _tiny_pipeline(tmp_path)
    lat = torch.randn(5, 4, 16, 16, generator=torch.Generator().manual_seed(9)).cuda()
    handle = pipe.wfc_loss.native_handle(cfg["out_channels"], lat.device)
    z1 = lat.clone().requires_grad_()
    l1 = GuidanceLossFunction.apply(z1, pipe.tile_vae._native, handle, 1 / 0.18215, 5.0)
    (g1,) = torch.autograd.grad(l1.sum(), z1)
    monkeypatch.setenv("WFD_MAX_BATCH", "2")
    pipe.tile_vae.refresh_native()
    z2 = lat.clone().requires_grad_()
    l2 = GuidanceLossFunction.apply(z2, pipe.tile_vae._native, handle, 1 / 0.18215, 5.0)
    (g2,) = torch.autograd.grad(l2.sum(), z2)
    assert _rel(l2.cpu(), l1.cpu()) < 1e-4 and _rel(g2.cpu(), g1.cpu()) < 5e-2


@pytest.mark.parametrize("arch,tileset,latent", [("32x32", "platform_32x32", 64), ("64x64", "install_64x64", 128),
                                                 ("64x64", "twilight_64x64", 64), ("64x64", "ashworld_64x64", 64),
                                                 ("64x64", "badlands_64x64", 64), ("64x64", "desert_64x64", 64),
                                                 ("64x64", "jungle_64x64", 64), ("64x64", "platform_64x64", 64),
                                                 ("64x64", "install_64x64", 64)])
def test_other_configs_against_the_oracle(arch, tileset, latent):
    """BASELINE config 5's architecture (DownDecoderBlock2D, 32x32 map from 64x64 latents), a 128x128 map, and the full
    architecture on EVERY tileset of config 3 (ice has its own golden tests): T_pad 1024 ... 4352 changes conv_out's
    channels per group, the N tiling of the WFC contraction and its K-chunk skip lists."""
    mx, my = load_wfc_mats(tileset)
    t_pad = round_up_channels(mx.shape[0])
    cfg = tilenet_config(arch, t_pad)
    torch.manual_seed(0)
    net = AutoencoderTile(**cfg).eval()
    lat = torch.randn(1, 4, latent, latent, generator=torch.Generator().manual_seed(2))
    want = orc.guidance_step(net.state_dict(), cfg, mx, my, lat, 5.0, torch.float32)
    net = net.cuda()
    wfc = WFCLossEinsum(mx, my).cuda()
    z = lat.cuda().requires_grad_()
    loss = GuidanceLossFunction.apply(z, net._native, wfc.native_handle(t_pad, z.device), 1 / 0.18215, 5.0)
    (grad,) = torch.autograd.grad(loss.sum(), z)
    print(arch, tileset, latent, "loss", loss.item(), want["loss"].item(), "grad rel", _rel(grad.cpu(), want["grad"]))
    assert _rel(loss.cpu(), want["loss"]) < 1e-3
    assert _rel(grad.cpu(), want["grad"]) < 5e-2


def test_img2img_from_a_tile_map(tmp_path):
    """img2img twin started from an integer tile map (image_is_wave=True): tile-VAE encode -> noised latents -> the same
    guided loop (reference pipeline_wfd_img2img.py:478-486, 667-700)."""
    from wavefunctiondiffusion_b200.wf_diffusers.pipeline_wfd_img2img import WaveFunctionDiffusionImg2ImgPipeline, preprocess_wave

    z = np.load(os.path.join(GOLD, "tiny_step.npz"))
    cfg = json.loads(str(z["cfg_json"]))
    T = int(z["T"])
    mats = np.unpackbits(z["mats"])[: 2 * T * T].reshape(2, T, T).astype(bool)
    torch.manual_seed(int(z["seed"]))
    net = AutoencoderTile(**cfg).eval()
    npz = str(tmp_path / "wfc.npz")
    np.savez(npz, wfc_mats=mats)
    pipe = WaveFunctionDiffusionImg2ImgPipeline(TinyImageVAE(), ConstantTextEncoder(), WhitespaceTokenizer(), net,
                                                RandomNoiseUNet(4, 16), AffineDDIMScheduler(), None, None, wfc_data_path=npz)
    pipe.set_progress_bar_config(disable=True)
    pipe.to("cuda")
    tiles = np.random.RandomState(0).randint(0, T, size=(16, 16))
    assert preprocess_wave(tiles, cfg["in_channels"]).shape == (1, cfg["in_channels"], 16, 16)
    out = pipe("a map", image=tiles, image_is_wave=True, strength=0.5, num_inference_steps=6, wfc_guidance_start_step=1,
               wfc_guidance_strength=5, wfc_guidance_final_steps=1, wfc_guidance_final_strength=10, output_type="np",
               generator=torch.Generator("cuda").manual_seed(0))
    assert out.waves.shape == (1, 16, 16, cfg["out_channels"]) and np.isfinite(out.waves).all()
    with pytest.raises(ValueError):
        pipe("x", image=tiles, image_is_wave=True, strength=1.5)
