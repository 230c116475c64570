"""Row-band sharding of ONE large map over several GPUs (BASELINE config 4: 256x256).

The reference decodes and scores a large map on one device (`pipeline_wfd.py:283-290` only widens `width`/`height`).
Here rank r of R owns rows [r*h/R, (r+1)*h/R) of the latent plane and the same rows of the wave; libwfd_b200 exchanges
what the single-device ops would have read across the cut (one halo row per 3x3 conv and for the vertical WFC pairs,
GroupNorm sums, the loss, attention keys/values and their gradients - include/wfd_b200.h, csrc/comm.cuh).

Transports: NCCL (one process per GPU under torchrun; the communicator is created from the libnccl that torch itself
loaded) and an in-process loopback group (R bands on one GPU, one host thread per band) for single-GPU tests.
"""
import ctypes as C
import os
import sys
import threading

import torch

from .. import _native as nat
from .native_ops import DecoderPlan, WfcHandle, _require_cuda, _z_dtype_code, compute_dtype16


def band_rows(h_total, rank, nranks):
    """[start, stop) rows of the plane that `rank` owns; bands are equal (the kernels tile whole 128-pixel patches)."""
    if h_total % nranks != 0:
        raise nat.WfdError(f"row bands: plane height {h_total} is not divisible by {nranks} ranks")
    hb = h_total // nranks
    return rank * hb, (rank + 1) * hb


def check_band_shape(h_total, w, nranks):
    """The local plane must still tile into 128-pixel patches (w_box * h_box == 128 with w_box | w, h_box | h/R)."""
    start, stop = band_rows(h_total, 0, nranks)
    hb = stop - start
    for wb in (128, 64, 32, 16, 8, 4, 2, 1):
        if w % wb == 0 and hb % (128 // wb) == 0:
            return hb
    raise nat.WfdError(f"row bands: a band of {hb}x{w} cells cannot be tiled into 128-cell patches")


def nccl_library_path():
    """Path of the libnccl this process has loaded (torch's bundled one), so the library joins the same NCCL build."""
    try:
        with open("/proc/self/maps") as f:
            for line in f:
                if "libnccl" in line:
                    return line.split()[-1]
    except OSError:
        pass
    try:
        import nvidia.nccl  # noqa: F401

        cand = os.path.join(os.path.dirname(nvidia.nccl.__file__), "lib", "libnccl.so.2")
        if os.path.exists(cand):
            return cand
    except Exception:
        pass
    return "libnccl.so.2"


class LoopbackGroup:
    """R bands on one GPU in one process, each driven by its own host thread (tests)."""

    def __init__(self, nranks):
        self.nranks = nranks
        self.handle = nat.lib().wfd_loop_group_create(nranks)
        if not self.handle:
            raise nat.WfdError("wfd_loop_group_create failed")
        self._slots = [None] * nranks
        self._barrier = threading.Barrier(nranks)

    def gather_rows(self, rank, t, dim):
        """Host-side counterpart of the all-gather: every thread leaves its band, all read the concatenation."""
        torch.cuda.current_stream().synchronize()
        self._slots[rank] = t
        self._barrier.wait(timeout=120)
        out = torch.cat([x.clone() for x in self._slots], dim=dim)
        torch.cuda.current_stream().synchronize()
        self._barrier.wait(timeout=120)
        return out

    def comm(self, rank):
        return BandComm._loopback(self, rank)

    def __del__(self):
        try:
            if self.handle:
                nat.lib().wfd_loop_group_destroy(C.c_void_p(self.handle))
                self.handle = None
        except Exception:
            pass


class BandComm:
    def __init__(self, handle, rank, nranks, keep=None):
        self.handle, self.rank, self.nranks, self._keep = handle, rank, nranks, keep

    def enable_peer_mailboxes(self, max_row_bytes=8 << 20):
        """Halo rows, GroupNorm sums and the loss as direct NVLink stores + flags instead of NCCL send/recv/all-reduce
        (include/wfd_b200.h, wfd_comm_peer_*). Every rank's mailbox handle travels over the default process group; a rank
        that cannot create or map a mailbox makes ALL ranks stay on NCCL (the transports must match)."""
        import torch.distributed as dist

        mine = C.create_string_buffer(64)
        ok = nat.lib().wfd_comm_peer_handle(self.handle, C.c_size_t(max_row_bytes), mine) == 0
        gathered = [None] * self.nranks
        dist.all_gather_object(gathered, mine.raw if ok else None)
        if any(g is None for g in gathered):
            return False
        blob = C.create_string_buffer(b"".join(gathered), 64 * self.nranks)
        attached = nat.lib().wfd_comm_peer_attach(self.handle, blob) == 0
        flags = [None] * self.nranks
        dist.all_gather_object(flags, attached)
        if not all(flags):
            raise nat.WfdError("row bands: peer mailboxes could be mapped on some ranks only; set WFD_BAND_P2P=0")
        self.peer = True
        if self.rank == 0 and os.environ.get("WFD_VERBOSE"):
            print("row bands: halo rows / GroupNorm sums / loss over peer mailboxes (%d ranks)" % self.nranks, file=sys.stderr)
        return True

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(128)
        nat.check(nat.lib().wfd_comm_unique_id(nccl_library_path().encode(), buf), "wfd_comm_unique_id")
        return buf.raw

    @staticmethod
    def nccl(rank, nranks, unique_id):
        h = C.c_void_p()
        nat.check(nat.lib().wfd_comm_create_nccl(nccl_library_path().encode(), C.c_char_p(unique_id), rank, nranks,
                                                 C.byref(h)), "wfd_comm_create_nccl")
        return BandComm(h, rank, nranks)

    @staticmethod
    def from_torch_distributed(device, max_row_bytes=8 << 20):
        """One band per torch.distributed rank; rank 0's NCCL id travels over the default process group. The small
        exchanges go through peer mailboxes over NVLink (rows of up to max_row_bytes; WFD_BAND_P2P=0 keeps them on NCCL)."""
        import torch.distributed as dist

        rank, nranks = dist.get_rank(), dist.get_world_size()
        box = [BandComm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        with torch.cuda.device(device):
            comm = BandComm.nccl(rank, nranks, box[0])
            if nranks > 1 and os.environ.get("WFD_BAND_P2P", "1") != "0":
                comm.enable_peer_mailboxes(max_row_bytes)
            return comm

    @staticmethod
    def _loopback(group, rank):
        h = C.c_void_p()
        nat.check(nat.lib().wfd_comm_create_loopback(C.c_void_p(group.handle), rank, C.byref(h)),
                  "wfd_comm_create_loopback")
        return BandComm(h, rank, group.nranks, keep=group)

    def gather_rows(self, t, dim):
        """Concatenate every rank's band along `dim`: torch.distributed all_gather between processes, a host-side
        exchange between the threads of a loopback group."""
        if self._keep is not None:
            return self._keep.gather_rows(self.rank, t, dim)
        return gather_rows(t, dim)

    def close(self):
        """Destroy the communicator. Call it after every BandGuidance that captured CUDA graphs over it was closed: NCCL
        keeps a communicator alive for the graphs that reference it, and tearing it down first can block."""
        if self.handle:
            nat.lib().wfd_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            import sys

            if sys.is_finalizing():
                return  # interpreter shutdown: destruction order is arbitrary, the process is going away anyway
            self.close()
        except Exception:
            pass


class BandGuidance:
    """One rank's share of the WFC-guidance step of one (h_total x w latent) map.

    step(z) takes the WHOLE latent plane (1, 4, h_total, w) - it is 4 channels, every rank keeps a copy - and returns the
    map's loss (identical on every rank) and d loss / d z for this rank's rows, (1, 4, h_total/R, w)."""

    def __init__(self, model, mat_x, mat_y, h_total, w, comm, device, dtype16=None):
        self.comm = comm
        self.rank, self.nranks = comm.rank, comm.nranks
        self.h_total, self.w = h_total, w
        self.h_band = check_band_shape(h_total, w, comm.nranks)
        self.row0 = self.rank * self.h_band
        self.device = device
        self.dtype16 = compute_dtype16() if dtype16 is None else dtype16
        lib = nat.lib()
        self.plan = DecoderPlan(model, h_total, w, 1, self.dtype16, device, band=(self.rank, self.nranks))
        nat.check(lib.wfd_decoder_attach_comm(self.plan.handle, comm.handle), "wfd_decoder_attach_comm")
        self.wfc = WfcHandle(mat_x, mat_y, model.config.out_channels, self.dtype16, device)
        self.H_total = self.plan.H * self.nranks  # map rows (the ice/platform decoders keep the latent resolution)
        nat.check(lib.wfd_wfc_attach_comm(self.wfc.handle, comm.handle, self.H_total), "wfd_wfc_attach_comm")
        self._graphs = {}

    def _check_z(self, z):
        _require_cuda(z, "row-band guidance")
        if tuple(z.shape) != (1, 4, self.h_total, self.w):
            raise nat.WfdError(f"row-band guidance: latents must be (1, 4, {self.h_total}, {self.w}), got {tuple(z.shape)}")
        return z.contiguous()

    def decode(self, z, in_scale=1.0):
        """Logits of this band: (1, T_pad, h_band, w) view (valid until the next call)."""
        z = self._check_z(z)
        self.plan.forward(z, in_scale)
        return self.plan.logits_view(1)

    def _launch(self, z, in_scale, strength, loss, dz):
        self.plan.generation += 1
        nat.check(nat.lib().wfd_guidance_step(self.plan.handle, self.wfc.handle, C.c_void_p(z.data_ptr()), _z_dtype_code(z),
                                              1, float(in_scale), float(strength),
                                              self.wfc.saved_ptr(1, self.plan.H, self.plan.W), C.c_void_p(loss.data_ptr()),
                                              C.c_void_p(dz.data_ptr()), nat.stream_ptr()), "wfd_guidance_step")

    def step(self, z, in_scale=1.0, strength=1.0, graph=False):
        """graph=True replays the step (kernels AND the NCCL exchanges between them) from a CUDA graph after two plain
        warm-up calls: at 8 bands the ~290 short launches per step are otherwise bound by the host. Every rank must make
        the same choice; the returned tensors are then static buffers that the next call overwrites."""
        z = self._check_z(z)
        if not graph:
            loss = torch.empty(1, dtype=torch.float32, device=self.device)
            dz = torch.empty((1, 4, self.h_band, self.w), dtype=torch.float32, device=self.device)
            self._launch(z, in_scale, strength, loss, dz)
            return loss, dz
        key = (z.dtype, float(in_scale), float(strength))
        g = self._graphs.get(key)
        if g is None:
            g = {"z": torch.zeros_like(z), "loss": torch.empty(1, dtype=torch.float32, device=self.device),
                 "dz": torch.empty((1, 4, self.h_band, self.w), dtype=torch.float32, device=self.device), "calls": 0,
                 "graph": None}
            self._graphs[key] = g
        g["z"].copy_(z)
        if g["graph"] is None and g["calls"] >= 2:
            torch.cuda.synchronize(self.device)
            graph_obj = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph_obj, capture_error_mode="thread_local"):
                self._launch(g["z"], in_scale, strength, g["loss"], g["dz"])
            g["graph"] = graph_obj
        if g["graph"] is not None:
            g["graph"].replay()
        else:
            self._launch(g["z"], in_scale, strength, g["loss"], g["dz"])
        g["calls"] += 1
        return g["loss"], g["dz"]

    def close(self):
        """Release the captured graphs (they hold the NCCL exchanges) before the communicator goes away."""
        self._graphs.clear()
        torch.cuda.synchronize(self.device)

    def wave(self):
        """(1, h_band, w, T_pad) wave of this band after step()."""
        return self.wfc.wave_view(1, self.plan.H, self.plan.W)


class BandGuidanceFunction(torch.autograd.Function):
    """strength * wfc_loss(softmax(decode(latents / 0.18215))) of ONE map computed in row bands: the loss is the map's
    loss on every rank, the gradient is gathered from the bands, so `torch.autograd.grad(loss, latents)` behaves exactly
    as with the single-device GuidanceLossFunction (reference pipeline_wfd.py:609-613)."""

    @staticmethod
    def forward(ctx, latents, bands, in_scale, strength, graph):
        loss, dz = bands.step(latents.detach(), in_scale, strength, graph=graph)
        ctx.save_for_backward(bands.comm.gather_rows(dz, 2))
        ctx.zdtype = latents.dtype
        return loss.clone()

    @staticmethod
    def backward(ctx, dloss):
        (dz,) = ctx.saved_tensors
        return (dz * dloss.float().view(-1, 1, 1, 1)).to(ctx.zdtype), None, None, None, None


def gather_rows(t, dim):
    """Concatenate every rank's band along `dim` (torch.distributed all_gather; bands are equal-sized)."""
    import torch.distributed as dist

    parts = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, t.contiguous())
    return torch.cat(parts, dim=dim)


def run_loopback(nranks, fn):
    """Run fn(rank, comm) for every band on its own thread and CUDA stream; returns the results in rank order."""
    group = LoopbackGroup(nranks)
    out = [None] * nranks
    err = [None] * nranks
    dev = torch.cuda.current_device()

    def work(r):
        try:
            torch.cuda.set_device(dev)
            with torch.cuda.stream(torch.cuda.Stream()):
                out[r] = fn(r, group.comm(r))
                torch.cuda.current_stream().synchronize()
        except BaseException as e:  # noqa: BLE001 - reported to the caller below
            err[r] = e

    threads = [threading.Thread(target=work, args=(r,)) for r in range(nranks)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in err:
        if e is not None:
            raise e
    return out
