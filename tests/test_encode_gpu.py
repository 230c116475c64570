"""GPU: AutoencoderTile.encode through libwfd_b200.so (the img2img wave input; reference wf_vae.py:68-148,
582-590) against the moments of the unmodified reference (tests/golden/tiny_encode.npz) and, at the ice tileset's full
size, against the CPU oracle. Tolerance: rel-L2 <= 2e-2 (bf16 operands, fp32 accumulation), as for the decoder's logits."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import wfd_oracle as orc
from wavefunctiondiffusion_b200.utils.tilesets import load_wfc_mats, round_up_channels, tilenet_config
from wavefunctiondiffusion_b200.wf_diffusers.wf_vae import AutoencoderTile

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    a, b = torch.as_tensor(np.asarray(a)).double(), torch.as_tensor(np.asarray(b)).double()
    return ((a - b).norm() / b.norm()).item()


def test_encode_small_against_reference_golden():
    z = np.load(os.path.join(GOLD, "tiny_encode.npz"))
    cfg = json.loads(str(z["cfg_json"]))
    torch.manual_seed(int(z["seed"]))
    net = AutoencoderTile(**cfg).eval().cuda()
    tiles = torch.from_numpy(z["tiles"]).cuda()
    onehot = torch.nn.functional.one_hot(tiles, num_classes=cfg["in_channels"]).permute(0, 3, 1, 2).float()
    with torch.no_grad():
        by_index = net.encode_tiles(tiles).latent_dist          # conv_in as a gather
        dense = net.encode(onehot).latent_dist                   # conv_in on the tensor cores
        soft = net.encode(torch.from_numpy(z["soft"]).cuda()).latent_dist
    assert by_index.parameters.shape == (2, 8, 16, 16) and by_index.mean.shape == (2, 4, 16, 16)
    assert _rel(by_index.parameters.cpu(), z["moments_onehot_f64"]) < 2e-2
    assert _rel(dense.parameters.cpu(), z["moments_onehot_f64"]) < 2e-2
    assert _rel(soft.parameters.cpu(), z["moments_soft_f64"]) < 2e-2
    # the two conv_in forms differ only in rounding (fp32 weights in the gather, 16-bit operands on the tensor cores):
    # they agree to the same 16-bit noise floor
    assert _rel(by_index.parameters.cpu(), dense.parameters.cpu()) < 2e-2
    # a (H, W) map is a batch of one; sampling works on the device
    one = net.encode_tiles(tiles[0]).latent_dist
    assert one.sample(torch.Generator("cuda").manual_seed(0)).shape == (1, 4, 16, 16)


def test_encode_ice_full_size_against_the_oracle():
    mx, _ = load_wfc_mats("ice_64x64")
    T = mx.shape[0]
    t_pad = round_up_channels(T)
    cfg = tilenet_config("64x64", t_pad)
    torch.manual_seed(0)
    net = AutoencoderTile(**cfg).eval()
    tiles = torch.from_numpy(np.random.RandomState(0).randint(0, T, size=(1, 64, 64)))
    onehot = torch.nn.functional.one_hot(tiles, num_classes=t_pad).permute(0, 3, 1, 2).float()
    want = orc.encode(net.state_dict(), cfg, onehot, torch.float32)
    net = net.cuda()
    with torch.no_grad():
        got = net.encode_tiles(tiles.cuda()).latent_dist.parameters
        got_dense = net.encode(onehot.cuda()).latent_dist.parameters
    print("ice encode: rel-L2 by index", _rel(got.cpu(), want), "dense", _rel(got_dense.cpu(), want))
    assert _rel(got.cpu(), want) < 2e-2 and _rel(got_dense.cpu(), want) < 2e-2
