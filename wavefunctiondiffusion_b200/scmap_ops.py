"""Device-side tile post-processing: what the reference's callers do on the host with `waves[0].argmax(axis=2)`
(wfd/webui/ui.py:228-243) - mirror symmetry (wfd/scmap/data_processing.py:191-196) and the id chain
gen_to_sc[shrink_to_gen[tile]] (wfd/scmap/map_display.py:55) - so that a (H, W) uint16 map leaves the GPU instead of the
(H, W, T_pad) fp32 wave (52 MB per 64x64 ice map). Integer gathers through libwfd_b200.so; bit-exact against numpy.
"""
import ctypes as C

import numpy as np
import torch

from . import _native as nat
from .wf_diffusers.native_ops import _require_cuda, argmax_tiles


class TileTables:
    """The three integer tables of one tileset on one device: `shrink_to_gen`, `gen_to_sc` (keys of the reference's
    wfc/*.npz, written by wfd/wfdf/datasets.py:130-146) and, optionally, the mirror `mapping` of tile_data/symmetry/*.npz."""

    def __init__(self, shrink_to_gen, gen_to_sc, symmetry_mapping=None, device="cuda"):
        def dev(a):
            return None if a is None else torch.as_tensor(np.asarray(a).astype(np.int32)).to(device).contiguous()

        self.shrink_to_gen, self.gen_to_sc, self.symmetry = dev(shrink_to_gen), dev(gen_to_sc), dev(symmetry_mapping)
        self.device = torch.device(device)

    @classmethod
    def from_npz(cls, wfc_data_path, symmetry_path=None, device="cuda"):
        z = np.load(wfc_data_path)
        s2g = z["shrink_to_gen"] if ("shrink_range" not in z.files or bool(z["shrink_range"])) else None
        mapping = np.load(symmetry_path)["mapping"] if symmetry_path else None
        return cls(s2g, z["gen_to_sc"], mapping, device)


def finish_tiles(tiles, tables, symmetry=False):
    """tiles: int64 (B, H, W) or (H, W) on the device -> (tiles', sc, bad): tiles' int64 (B, H, W') shrink ids after the
    optional mirror, sc uint16 (B, H, W') StarCraft ids, bad = number of ids outside a table (they come out as 0xFFFF)."""
    _require_cuda(tiles, "finish_tiles")
    t = tiles if tiles.dim() == 3 else tiles.unsqueeze(0)
    t = t.to(torch.int64).contiguous()
    B, H, W = t.shape
    symm = tables.symmetry if symmetry else None
    if symmetry and symm is None:
        raise ValueError("finish_tiles: symmetry requested but the tables carry no mirror mapping")
    Wo = 2 * W if symmetry else W
    tiles_out = torch.empty((B, H, Wo), dtype=torch.int64, device=t.device)
    sc = torch.empty((B, H, Wo), dtype=torch.uint16, device=t.device)
    bad = torch.zeros(1, dtype=torch.int64, device=t.device)
    s2g = tables.shrink_to_gen

    def p(x):
        return None if x is None else C.c_void_p(x.data_ptr())

    nat.check(nat.lib().wfd_tiles_finish(p(t), B, H, W, p(symm), 0 if symm is None else symm.numel(), p(s2g),
                                         0 if s2g is None else s2g.numel(), p(tables.gen_to_sc), tables.gen_to_sc.numel(),
                                         p(tiles_out), p(sc), p(bad), nat.stream_ptr()), "wfd_tiles_finish")
    return tiles_out, sc, bad


def wave_to_map(wave, tables, count=None, symmetry=False):
    """wave (B, H, W, T_pad) on the device (fp32 / fp16 / bf16) -> (tiles', sc, bad) without a host round trip:
    numpy.argmax semantics over the first `count` channels (all of them when None), then finish_tiles."""
    if count is not None and count < wave.shape[-1]:
        wave = wave[..., :count].contiguous()
    return finish_tiles(argmax_tiles(wave), tables, symmetry)
