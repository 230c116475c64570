"""Worker of tests/test_rowband_gpu.py::test_band_plan_over_nccl (one process per GPU under torchrun): the row-band
guidance step over the NCCL transport against the single plan — logits bit-exact, loss and gradient within the fp32
summation order of the key/value gradients. Prints NCCL_BAND_OK on rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from wavefunctiondiffusion_b200.wf_diffusers.native_ops import GuidanceLossFunction  # noqa: E402
from wavefunctiondiffusion_b200.wf_diffusers.rowband import BandComm, BandGuidance, gather_rows  # noqa: E402
from wavefunctiondiffusion_b200.wf_diffusers.wf_vae import AutoencoderTile  # noqa: E402
from wavefunctiondiffusion_b200.wfdf.losses import WFCLossEinsum  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    z = np.load(os.path.join(ROOT, "tests", "golden", "tiny_step.npz"))
    cfg = json.loads(str(z["cfg_json"]))
    T = int(z["T"])
    mats = np.unpackbits(z["mats"])[: 2 * T * T].reshape(2, T, T).astype(bool)
    torch.manual_seed(int(z["seed"]))
    net = AutoencoderTile(**cfg).eval().to(dev)
    lat = (torch.randn(1, 4, 32, 32, generator=torch.Generator().manual_seed(11)) * 0.8).to(dev)
    comm = BandComm.from_torch_distributed(dev)
    bg = BandGuidance(net, mats[0], mats[1], 32, 32, comm, dev)
    for use_graph in (False, True):
        for _ in range(4 if use_graph else 1):  # graph=True captures on the third call
            loss, dz = bg.step(lat, 1.0 / 0.18215, 5.0, graph=use_graph)
        full_dz = gather_rows(dz.clone(), 2)
        logits = gather_rows(bg.plan.logits_view(1).float().contiguous(), 2)
        losses = [torch.empty_like(loss) for _ in range(world)]
        dist.all_gather(losses, loss.clone())
        if rank == 0:
            wfc = WFCLossEinsum(mats[0], mats[1]).to(dev)
            x = lat.clone().requires_grad_()
            want = GuidanceLossFunction.apply(x, net._native, wfc.native_handle(cfg["out_channels"], dev), 1.0 / 0.18215, 5.0)
            (g,) = torch.autograd.grad(want.sum(), x)
            ref_logits = net._native.plan(1, 32, 32, dev).logits_view(1).float()
            assert all(torch.equal(l, losses[0]) for l in losses), "every band must hold the same loss"
            assert torch.equal(logits, ref_logits), "band logits over NCCL must be bit-exact"
            assert abs(loss.item() - want.item()) <= 1e-6 * abs(want.item())
            rel = ((full_dz - g).norm() / g.norm()).item()
            assert rel < 2e-3, rel
            print("graph=%s loss %.7f (single %.7f) grad rel %.2e" % (use_graph, loss.item(), want.item(), rel))
    bg.close()
    comm.close()
    dist.barrier()
    if rank == 0:
        print("NCCL_BAND_OK ranks=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
