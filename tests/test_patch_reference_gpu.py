"""GPU: the shipped shim of INTEGRATION.md section B (wavefunctiondiffusion_b200/patch_reference.py) applied to the
reference's OWN objects. The unmodified reference modules are imported through oracle/reference_loader.py from the
authoring container's /root/reference or from the offline install under baseline/_ref (git-ignored; it travels to the
GPU box with the snapshot); the test is skipped where neither exists. The reference's eager PyTorch path on the same
B200 is the comparison: same lines as pipeline_wfd.py:609-613."""
import json
import os

import numpy as np
import pytest
import torch

from oracle.reference_loader import load_reference, reference_root
from wavefunctiondiffusion_b200.patch_reference import patch_autoencoder_tile, patch_wfc_loss, unpatch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(reference_root() is None, reason="the reference is not installed")]
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    a, b = torch.as_tensor(np.asarray(a)).double(), torch.as_tensor(np.asarray(b)).double()
    return ((a - b).norm() / b.norm()).item()


def _guidance_lines(net, wfc_loss, latents, strength):
    """pipeline_wfd.py:609-613 (and decode_latents_tile :402-405) as the reference writes them."""
    with torch.enable_grad():
        _latents = latents.detach().requires_grad_()
        image = net.decode(1 / 0.18215 * _latents).sample
        wave = image.softmax(axis=1)
        _loss = wfc_loss(wave) * strength
        (grad,) = torch.autograd.grad(_loss.sum(), _latents)
    return image.detach().float().cpu(), _loss.detach().float().cpu(), grad.float().cpu()


@pytest.mark.parametrize("name", ["tiny_step.npz", "tiny_down_step.npz"])
def test_reference_objects_on_the_cuda_path(name):
    vae, losses = load_reference()
    z = np.load(os.path.join(GOLD, name))
    cfg = json.loads(str(z["cfg_json"]))
    T = int(z["T"])
    mats = np.unpackbits(z["mats"])[: 2 * T * T].reshape(2, T, T).astype(bool)
    torch.manual_seed(int(z["seed"]))
    net = vae.AutoencoderTile(**cfg).eval().cuda()
    wfc_loss = losses.WFCLossEinsum(mats[0], mats[1]).cuda()
    lat = torch.from_numpy(z["latents"]).cuda()
    eager = _guidance_lines(net, wfc_loss, lat, 5.0)           # the reference's eager fp32 path on this GPU
    patch_autoencoder_tile(net)
    patch_wfc_loss(wfc_loss)
    native = _guidance_lines(net, wfc_loss, lat, 5.0)          # same objects, same lines, libwfd_b200.so underneath
    assert type(net.decode(lat).sample) is torch.Tensor and type(net.decode(lat)).__name__ == "DecoderOutput"
    for got, want in ((native, eager), (native, (z["logits_f64"], z["loss_f64"], z["grad_f64"]))):
        assert _rel(got[0], want[0]) < 2e-2
        assert _rel(got[1], want[1]) < 1e-3
        assert _rel(got[2], want[2]) < 5e-2
    # precision switch of the reference pipeline (set_precision: .half() on the modules): packed weights are rebuilt
    net.half()
    wfc_loss.half()
    half = _guidance_lines(net, wfc_loss, lat.half(), 5.0)
    assert _rel(half[1], z["loss_f64"]) < 5e-3
    net.float()
    wfc_loss.float()
    unpatch(net)
    unpatch(wfc_loss)
    again = _guidance_lines(net, wfc_loss, lat, 5.0)
    # (eager fp32 convolutions on this GPU go through cuDNN's TF32 / non-deterministic algorithms: ~1e-3 run to run)
    assert _rel(again[1], eager[1]) < 1e-4 and _rel(again[2], eager[2]) < 5e-3
