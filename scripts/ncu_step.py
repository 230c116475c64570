"""Two guidance steps at BASELINE config 2 (ice 64x64, batch 8) for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wavefunctiondiffusion_b200.utils.tilesets import load_wfc_mats, round_up_channels, tilenet_config
from wavefunctiondiffusion_b200.wf_diffusers.native_ops import GuidanceLossFunction
from wavefunctiondiffusion_b200.wf_diffusers.wf_vae import AutoencoderTile
from wavefunctiondiffusion_b200.wfdf.losses import WFCLossEinsum

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mx, my = load_wfc_mats("ice_64x64")
t_pad = round_up_channels(mx.shape[0])
torch.manual_seed(0)
net = AutoencoderTile(**tilenet_config("64x64", t_pad)).eval().cuda()
wfc = WFCLossEinsum(mx, my).cuda()
lat = torch.randn(B, 4, 64, 64, generator=torch.Generator().manual_seed(1234)).cuda()
h = wfc.native_handle(t_pad, lat.device)
for _ in range(2):
    loss = GuidanceLossFunction.apply(lat, net._native, h, 1 / 0.18215, 5.0)
torch.cuda.synchronize()
print("loss", loss.tolist()[:2])
