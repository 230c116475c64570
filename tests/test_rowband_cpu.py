"""CPU, world_size 2 over gloo: the exchange contract of the row-band plan, stated with torch ops.

The CUDA library performs these exchanges itself (csrc/comm.cu); this test pins down on the CPU WHAT has to be exchanged
for a banded run to equal the single-device ops of the reference: one halo row per 3x3 conv (resnet.py:416) and for the
vertical WFC pairs (losses.py:60-66), all-reduced GroupNorm sums (resnet.py:407), gathered attention keys/values
(attention.py:329-380), an all-reduced loss (losses.py:69-75) - and exercises the host helpers of rowband.py.
"""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from oracle import wfd_oracle as orc
from wavefunctiondiffusion_b200 import _native as nat
from wavefunctiondiffusion_b200.wf_diffusers.rowband import band_rows, check_band_shape, gather_rows


def test_band_rows_partition_the_plane():
    for h, R in ((64, 1), (64, 2), (64, 8), (256, 8)):
        rows = []
        for r in range(R):
            lo, hi = band_rows(h, r, R)
            rows += list(range(lo, hi))
        assert rows == list(range(h))
    with pytest.raises(nat.WfdError):
        band_rows(30, 0, 4)


def test_band_shape_must_tile_into_128_cell_patches():
    assert check_band_shape(256, 256, 8) == 32
    assert check_band_shape(64, 64, 8) == 8
    with pytest.raises(nat.WfdError):
        check_band_shape(64, 64, 64)  # one row of 64 cells is not a whole patch


def _halo(x):
    """x: (1, C, h_band, W) -> (1, C, h_band + 2, W) with the neighbour bands' edge rows (zeros at the map border)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    top = torch.zeros_like(x[:, :, :1])
    bottom = torch.zeros_like(x[:, :, :1])
    ops = []
    first, last = x[:, :, :1].contiguous(), x[:, :, -1:].contiguous()
    if rank > 0:
        ops += [dist.P2POp(dist.isend, first, rank - 1), dist.P2POp(dist.irecv, top, rank - 1)]
    if rank < world - 1:
        ops += [dist.P2POp(dist.isend, last, rank + 1), dist.P2POp(dist.irecv, bottom, rank + 1)]
    for req in dist.batch_isend_irecv(ops) if ops else []:
        req.wait()
    return torch.cat([top, x, bottom], dim=2)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(3)
        C, H, W, T = 16, 8, 8, 6
        x = torch.randn(1, C, H, W, generator=g, dtype=torch.float64)
        wconv = torch.randn(C, C, 3, 3, generator=g, dtype=torch.float64)
        lo, hi = band_rows(H, rank, world)
        xb = x[:, :, lo:hi]
        out = {}
        # 3x3 conv: one halo row, zero padding only along x and at the map border
        yb = F.conv2d(_halo(xb), wconv, padding=(0, 1))
        out["conv"] = (gather_rows(yb, 2) - F.conv2d(x, wconv, padding=1)).abs().max().item()
        # GroupNorm: per-group {sum, sum of squares} all-reduced, count of the whole plane
        G = 4
        xg = xb.reshape(1, G, -1)
        sums = torch.stack([xg.sum(-1), (xg * xg).sum(-1)])
        dist.all_reduce(sums)
        n = (C // G) * H * W
        mean = sums[0] / n
        rstd = 1.0 / torch.sqrt(sums[1] / n - mean * mean + 1e-6)
        gnb = ((xb.reshape(1, G, -1) - mean[..., None]) * rstd[..., None]).reshape(xb.shape)
        out["gn"] = (gather_rows(gnb, 2) - F.group_norm(x, G, eps=1e-6)).abs().max().item()
        # attention: local queries against gathered keys / values
        tok = lambda t: t.permute(0, 2, 3, 1).reshape(1, -1, C)  # noqa: E731
        qb = tok(xb)
        kv = gather_rows(tok(xb), 1)
        ob = torch.softmax(qb @ kv.transpose(1, 2) / C ** 0.5, -1) @ kv
        full = torch.softmax(tok(x) @ tok(x).transpose(1, 2) / C ** 0.5, -1) @ tok(x)
        out["attn"] = (gather_rows(ob, 1) - full).abs().max().item()
        # WFC loss: vertical pairs across the cut read the halo row of the wave; partial sums all-reduced
        mats = (torch.rand(2, T, T, generator=g) > 0.5).numpy()
        wave = torch.softmax(torch.randn(1, T, H, W, generator=g, dtype=torch.float64), 1)
        wb = _halo(wave[:, :, lo:hi])
        nx, ny = torch.from_numpy(~mats[0]).double(), torch.from_numpy(~mats[1]).double()
        own = wb[:, :, 1:-1]
        part = torch.einsum("bihw,ij,bjhw->", own[..., :-1], nx, own[..., 1:])          # horizontal pairs of my rows
        part = part + torch.einsum("bihw,ij,bjhw->", wb[:, :, 1:-1], ny, wb[:, :, 2:])  # my rows with the row below
        if rank == world - 1:  # the bottom halo of the last band is outside the map: it is zero, so nothing was added
            assert wb[:, :, -1].abs().max().item() == 0.0
        tot = part.reshape(1).clone()
        dist.all_reduce(tot)
        P = H * (W - 1) + (H - 1) * W
        want = orc.wfc_loss(wave, mats[0], mats[1])
        out["loss"] = abs(tot.item() / P - float(np.asarray(want).reshape(-1)[0]))
        q.put((rank, out))
    except BaseException as e:  # noqa: BLE001 - reported through the queue so the parent does not wait for a timeout
        q.put((rank, {"error: " + repr(e): 1.0}))
        raise
    finally:
        dist.destroy_process_group()


def test_gloo_world2_band_exchanges_reproduce_the_single_device_ops():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29850 + os.getpid() % 100
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, out in res:
        for k, v in out.items():
            assert v < 1e-9, (rank, k, v)
