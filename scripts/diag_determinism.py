"""Diagnostic: run-to-run and sample-to-sample reproducibility of the CUDA path, plus a first timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wavefunctiondiffusion_b200.utils.tilesets import load_wfc_mats, round_up_channels, tilenet_config
from wavefunctiondiffusion_b200.wf_diffusers.native_ops import GuidanceLossFunction
from wavefunctiondiffusion_b200.wf_diffusers.wf_vae import AutoencoderTile
from wavefunctiondiffusion_b200.wfdf.losses import WFCLossEinsum

def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()

mx, my = load_wfc_mats("ice_64x64")
t_pad = round_up_channels(mx.shape[0])
torch.manual_seed(0)
net = AutoencoderTile(**tilenet_config("64x64", t_pad)).eval().cuda()
wfc = WFCLossEinsum(mx, my).cuda()
dev = torch.device("cuda:0")
handle = wfc.native_handle(t_pad, dev)
g = torch.Generator().manual_seed(5)
lat = torch.randn(8, 4, 64, 64, generator=g).cuda()

def step(z):
    z = z.clone().requires_grad_()
    loss = GuidanceLossFunction.apply(z, net._native, handle, 1 / 0.18215, 5.0)
    (grad,) = torch.autograd.grad(loss.sum(), z)
    plan = net._native.plan(z.shape[0], 64, 64, dev)
    return loss.detach().clone(), grad.clone(), plan.logits_view(z.shape[0]).float().clone()

l1, g1, lg1 = step(lat[:1])
l2, g2, lg2 = step(lat[:1])
print("run-to-run B=1: loss", abs(l1 - l2).item(), "logits rel", rel(lg1, lg2), "grad rel", rel(g1, g2))
two = torch.cat([lat[:1], lat[:1]])
l3, g3, lg3 = step(two)
print("same sample twice in a batch: logits rel", rel(lg3[0], lg3[1]), "grad rel", rel(g3[0], g3[1]), "vs single", rel(g3[0], g1[0]))
for B in (1, 8):
    z = lat[:B]
    step(z); torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(3):
        step(z)
    torch.cuda.synchronize()
    print("B=%d: %.2f ms/step" % (B, (time.time() - t0) / 3 * 1e3))
