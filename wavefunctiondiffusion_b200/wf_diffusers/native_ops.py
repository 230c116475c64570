"""torch <-> libwfd_b200.so glue for the tile-VAE decoder and the WFC loss: plan/handle caches, workspace ownership and
the autograd Functions that keep `torch.autograd.grad(loss, latents)` (reference pipeline_wfd.py:613) working.

torch is used for device memory, streams and autograd bookkeeping only; all arithmetic happens in the library.
"""
import ctypes as C
import os

import numpy as np
import torch

from .. import _native as nat

_DT16 = {torch.float16: 0, torch.bfloat16: 1}


def compute_dtype16():
    """Operand/activation type of the kernels: bf16 (default) or fp16 via WFD_DTYPE16=fp16."""
    return 0 if os.environ.get("WFD_DTYPE16", "bf16").lower() in ("fp16", "half", "float16") else 1


def _torch16(dtype16):
    return torch.bfloat16 if dtype16 else torch.float16


def _require_cuda(t, what):
    if not t.is_cuda:
        raise nat.WfdError(
            f"{what}: tensor is on {t.device}; wavefunctiondiffusion_b200 has no CPU path (move the model and inputs "
            "to a CUDA device)"
        )


def _z_dtype_code(z):
    if z.dtype == torch.float32:
        return 0
    if z.dtype == torch.float16:
        return 1
    raise nat.WfdError(f"latents must be float32 or float16, got {z.dtype}")


class DecoderPlan:
    """One wfd_decoder handle + its workspace for a fixed (h, w, B_max, dtype16).

    band=(rank, nranks) builds a row-band plan instead (see rowband.py): h is then the height of the whole latent
    plane, the plan decodes rows [rank*h/nranks, (rank+1)*h/nranks) of ONE map and needs a communicator attached."""

    def __init__(self, model, h, w, b_max, dtype16, device, band=None):
        lib = nat.lib()
        cfg = model.config
        c = nat.DecoderConfig()
        c.latent_channels = cfg.latent_channels
        c.out_channels = cfg.out_channels
        nb = len(cfg.block_out_channels)
        c.n_blocks = nb
        for i in range(nb):
            c.block_out_channels[i] = cfg.block_out_channels[i]
            c.conv_groups[i] = cfg.conv_groups[i]
            c.block_down[i] = 1 if cfg.up_block_types[i] == "DownDecoderBlock2D" else 0
        c.conv_init_groups = cfg.conv_init_groups
        c.layers_per_block = cfg.layers_per_block
        c.norm_num_groups = cfg.norm_num_groups
        self.handle = C.c_void_p()
        self.device = device
        self.h, self.w, self.b_max, self.dtype16 = h, w, b_max, dtype16
        self.t_pad = cfg.out_channels
        with torch.cuda.device(device):
            if band is None:
                nat.check(lib.wfd_decoder_create(C.byref(c), h, w, b_max, dtype16, C.byref(self.handle)),
                          "wfd_decoder_create")
            else:
                if b_max != 1:
                    raise nat.WfdError("row-band plans decode one map (batch 1)")
                nat.check(lib.wfd_decoder_create_band(C.byref(c), h, w, band[0], band[1], dtype16, C.byref(self.handle)),
                          "wfd_decoder_create_band")
                self.h_total = h
                self.h = h // band[1]
            self.band = band
            for key, t in model.state_dict().items():
                if not (key.startswith("decoder.") or key.startswith("post_quant_conv.")):
                    continue
                host = t.detach().to("cpu", torch.float32).contiguous()
                nat.check(lib.wfd_decoder_set_weight(self.handle, key.encode(), C.c_void_p(host.data_ptr()), host.numel()),
                          "wfd_decoder_set_weight(%s)" % key)
            nat.check(lib.wfd_decoder_finalize_weights(self.handle), "wfd_decoder_finalize_weights")
            nbytes = lib.wfd_decoder_workspace_bytes(self.handle)
            self.workspace = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
            base = (self.workspace.data_ptr() + 255) // 256 * 256
            nat.check(lib.wfd_decoder_bind_workspace(self.handle, C.c_void_p(base), nbytes), "wfd_decoder_bind_workspace")
        H, W = C.c_int(), C.c_int()
        lib.wfd_decoder_out_hw(self.handle, C.byref(H), C.byref(W))
        self.H, self.W = H.value, W.value
        self.generation = 0  # bumped by every forward: backward must belong to the latest one

    def __del__(self):
        try:
            if self.handle:
                nat.lib().wfd_decoder_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def forward(self, z, in_scale=1.0):
        _require_cuda(z, "decode")
        z = z.contiguous()
        self.generation += 1
        nat.check(nat.lib().wfd_decoder_forward(self.handle, C.c_void_p(z.data_ptr()), _z_dtype_code(z), z.shape[0],
                                                float(in_scale), nat.stream_ptr()), "wfd_decoder_forward")

    def logits_view(self, B):
        """Zero-copy (B, T_pad, H, W) channels-last view of the internal logits buffer (valid until the next forward)."""
        ptr = nat.lib().wfd_decoder_logits(self.handle)
        off = ptr - self.workspace.data_ptr()
        n = B * self.H * self.W * self.t_pad
        flat = self.workspace[off:off + 2 * n].view(_torch16(self.dtype16))
        return flat.view(B, self.H, self.W, self.t_pad).permute(0, 3, 1, 2)

    def export_logits(self, B, dtype):
        """(B, T_pad, H, W) tensor with channels-last strides, in `dtype` (fp32, or the plan's 16-bit type)."""
        out = torch.empty((B, self.t_pad, self.H, self.W), dtype=dtype, device=self.device,
                          memory_format=torch.channels_last)
        code = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}[dtype]
        nat.check(nat.lib().wfd_decoder_export_logits(self.handle, B, C.c_void_p(out.data_ptr()), code, 1,
                                                      nat.stream_ptr()), "wfd_decoder_export_logits")
        return out

    def backward(self, dlogits, B):
        """dlogits: (B, T_pad, H, W) any float dtype/layout, or None to use the internal dlogits buffer."""
        lib = nat.lib()
        if dlogits is not None:
            if dlogits.is_contiguous(memory_format=torch.channels_last) and dlogits.dtype in (torch.float32, _torch16(self.dtype16)):
                code = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}[dlogits.dtype]
                layout = 1
                src = dlogits
            else:
                src = dlogits.contiguous()
                if src.dtype not in (torch.float32, torch.float16):
                    src = src.float()
                code = 0 if src.dtype == torch.float32 else 1
                layout = 0
            nat.check(lib.wfd_decoder_import_dlogits(self.handle, B, C.c_void_p(src.data_ptr()), code, layout,
                                                     nat.stream_ptr()), "wfd_decoder_import_dlogits")
        dz = torch.empty((B, 4, self.h, self.w), dtype=torch.float32, device=self.device)
        nat.check(lib.wfd_decoder_backward(self.handle, None, B, None, C.c_void_p(dz.data_ptr()), nat.stream_ptr()),
                  "wfd_decoder_backward")
        return dz


class NativeDecoder:
    """Per-model cache of DecoderPlans; rebuilt when the weights change (load_state_dict / .to / .half)."""

    def __init__(self, model):
        import weakref

        self._model = weakref.ref(model)
        self._plans = {}

    def invalidate(self):
        self._plans.clear()

    def plan(self, B, h, w, device):
        dtype16 = compute_dtype16()
        key = (h, w, dtype16, str(device))
        p = self._plans.get(key)
        if p is None or p.b_max < B:
            self._plans.pop(key, None)
            p = DecoderPlan(self._model(), h, w, B, dtype16, device)
            self._plans[key] = p
        return p


class EncoderPlan:
    """One wfd_encoder handle + workspace for a fixed (H, W, B_max, dtype16): AutoencoderTile.encode on the device
    (reference wf_vae.py:582-590). Forward only."""

    def __init__(self, model, H, W, b_max, dtype16, device):
        lib = nat.lib()
        cfg = model.config
        c = nat.DecoderConfig()
        c.latent_channels = cfg.latent_channels
        c.out_channels = cfg.in_channels
        nb = len(cfg.block_out_channels)
        c.n_blocks = nb
        for i in range(nb):
            c.block_out_channels[i] = cfg.block_out_channels[i]
            c.conv_groups[i] = cfg.conv_groups[i]
            c.block_down[i] = 0 if cfg.down_block_types[i] == "EvenEncoderBlock2D" else 1
        c.conv_init_groups = cfg.conv_init_groups
        c.layers_per_block = cfg.layers_per_block
        c.norm_num_groups = cfg.norm_num_groups
        self.handle = C.c_void_p()
        self.device, self.H, self.W, self.b_max, self.dtype16 = device, H, W, b_max, dtype16
        self.n_moments = 2 * cfg.latent_channels
        with torch.cuda.device(device):
            nat.check(lib.wfd_encoder_create(C.byref(c), H, W, b_max, dtype16, C.byref(self.handle)), "wfd_encoder_create")
            for key, t in model.state_dict().items():
                if not (key.startswith("encoder.") or key.startswith("quant_conv.")):
                    continue
                host = t.detach().to("cpu", torch.float32).contiguous()
                nat.check(lib.wfd_encoder_set_weight(self.handle, key.encode(), C.c_void_p(host.data_ptr()), host.numel()),
                          "wfd_encoder_set_weight(%s)" % key)
            nat.check(lib.wfd_encoder_finalize_weights(self.handle), "wfd_encoder_finalize_weights")
            nbytes = lib.wfd_encoder_workspace_bytes(self.handle)
            self.workspace = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
            base = (self.workspace.data_ptr() + 255) // 256 * 256
            nat.check(lib.wfd_encoder_bind_workspace(self.handle, C.c_void_p(base), nbytes), "wfd_encoder_bind_workspace")

    def __del__(self):
        try:
            if self.handle:
                nat.lib().wfd_encoder_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def encode_wave(self, x):
        """x: (B, T_pad, H, W) fp32 / fp16 on the device -> moments fp32 (B, 2 * latent, H, W)."""
        x = x.contiguous()
        if x.dtype not in (torch.float32, torch.float16):
            x = x.float()
        out = torch.empty((x.shape[0], self.n_moments, self.H, self.W), dtype=torch.float32, device=self.device)
        nat.check(nat.lib().wfd_encoder_forward_wave(self.handle, C.c_void_p(x.data_ptr()), 0 if x.dtype == torch.float32 else 1,
                                                     x.shape[0], C.c_void_p(out.data_ptr()), nat.stream_ptr()),
                  "wfd_encoder_forward_wave")
        return out

    def encode_tiles(self, tiles):
        """tiles: int64 (B, H, W) on the device (the one-hot wave by its indices) -> moments fp32."""
        t = tiles.to(torch.int64).contiguous()
        out = torch.empty((t.shape[0], self.n_moments, self.H, self.W), dtype=torch.float32, device=self.device)
        nat.check(nat.lib().wfd_encoder_forward_tiles(self.handle, C.c_void_p(t.data_ptr()), t.shape[0],
                                                      C.c_void_p(out.data_ptr()), nat.stream_ptr()), "wfd_encoder_forward_tiles")
        return out


class NativeEncoder:
    """Per-model cache of EncoderPlans (rebuilt when the weights change)."""

    def __init__(self, model):
        import weakref

        self._model = weakref.ref(model)
        self._plans = {}

    def invalidate(self):
        self._plans.clear()

    @staticmethod
    def supported(config):
        return all(k == "EvenEncoderBlock2D" for k in config.down_block_types)

    def plan(self, B, H, W, device):
        dtype16 = compute_dtype16()
        key = (H, W, dtype16, str(device))
        p = self._plans.get(key)
        if p is None or p.b_max < B:
            self._plans.pop(key, None)
            p = EncoderPlan(self._model(), H, W, B, dtype16, device)
            self._plans[key] = p
        return p


class TileDecodeFunction(torch.autograd.Function):
    """logits = AutoencoderTile._decode(z) (reference wf_vae.py:592-599) with the data gradient wired to autograd."""

    @staticmethod
    def forward(ctx, z, native):
        _require_cuda(z, "AutoencoderTile.decode")
        B, _, h, w = z.shape
        plan = native.plan(B, h, w, z.device)
        plan.forward(z)
        ctx.plan, ctx.B, ctx.zdtype, ctx.generation = plan, B, z.dtype, plan.generation
        out_dtype = z.dtype if z.dtype in (torch.float32, _torch16(plan.dtype16)) else torch.float32
        out = plan.export_logits(B, out_dtype)
        return out.to(z.dtype) if out.dtype != z.dtype else out

    @staticmethod
    def backward(ctx, dlogits):
        if ctx.generation != ctx.plan.generation:
            raise nat.WfdError(
                "AutoencoderTile.decode: a later decode reused this model's activation workspace before backward of "
                "an earlier one ran; keep one outstanding decode per model (the guidance loop does)"
            )
        dz = ctx.plan.backward(dlogits, ctx.B)
        return dz.to(ctx.zdtype), None


# ------------------------------------------------------------------------------------------------ WFC loss
class WfcHandle:
    """wfd_wfc handle for one tileset (adjacency matrices) and 16-bit type; owns the `saved` block of the last forward."""

    def __init__(self, mat_x, mat_y, t_pad, dtype16, device):
        mx = np.ascontiguousarray(np.asarray(mat_x) != 0, dtype=np.uint8)
        my = np.ascontiguousarray(np.asarray(mat_y) != 0, dtype=np.uint8)
        assert mx.shape[0] == mx.shape[1] and mx.shape == my.shape, "WFC matrix size mismatch"
        self.T, self.t_pad, self.dtype16, self.device = mx.shape[0], t_pad, dtype16, device
        self.handle = C.c_void_p()
        with torch.cuda.device(device):
            nat.check(nat.lib().wfd_wfc_create(mx.ctypes.data_as(C.c_void_p), my.ctypes.data_as(C.c_void_p), self.T, t_pad,
                                               dtype16, C.byref(self.handle)), "wfd_wfc_create")
        self._saved = None
        self._saved_ptr = None

    def __del__(self):
        try:
            if self.handle:
                nat.lib().wfd_wfc_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def set_sparse(self, on):
        """Alternative contraction: CUDA-core gather over the allowed pairs instead of the dense tensor-core
        product; same results to rounding. Off by default."""
        with torch.cuda.device(self.device):
            nat.check(nat.lib().wfd_wfc_set_sparse(self.handle, int(bool(on))), "wfd_wfc_set_sparse")

    def saved_ptr(self, B, H, W):
        n = nat.lib().wfd_wfc_saved_bytes(self.handle, B, H, W)
        if self._saved is None or self._saved.numel() < n + 256:
            self._saved = torch.empty(n + 256, dtype=torch.uint8, device=self.device)
        self._saved_ptr = (self._saved.data_ptr() + 255) // 256 * 256
        return C.c_void_p(self._saved_ptr)

    def new_saved(self, B, H, W):
        """A fresh `saved` block for ONE forward, owned by the caller (the autograd context): several forwards of the
        same handle may then be outstanding before their backwards run. Returns (tensor, aligned pointer)."""
        n = nat.lib().wfd_wfc_saved_bytes(self.handle, B, H, W)
        t = torch.empty(n + 256, dtype=torch.uint8, device=self.device)
        return t, (t.data_ptr() + 255) // 256 * 256

    def forward(self, in_ptr, is_logits, B, H, W, strength, saved_ptr=None):
        """saved_ptr=None uses (and overwrites) the handle's own block - the fused guidance path and wave_view()."""
        loss = torch.empty(B, dtype=torch.float32, device=self.device)
        saved = self.saved_ptr(B, H, W) if saved_ptr is None else C.c_void_p(saved_ptr)
        nat.check(nat.lib().wfd_wfc_forward(self.handle, C.c_void_p(in_ptr), int(is_logits), B, H, W, float(strength),
                                            saved, C.c_void_p(loss.data_ptr()), nat.stream_ptr()),
                  "wfd_wfc_forward")
        return loss

    def backward(self, dloss, out_ptr, saved_ptr=None):
        nat.check(nat.lib().wfd_wfc_backward(self.handle, C.c_void_p(self._saved_ptr if saved_ptr is None else saved_ptr),
                                             None if dloss is None else C.c_void_p(dloss.data_ptr()),
                                             C.c_void_p(out_ptr), nat.stream_ptr()), "wfd_wfc_backward")

    def count_violations(self, tiles):
        """tiles: int64 (B, H, W) or (H, W) on the device -> int64 (B, 2): forbidden horizontal / vertical pairs."""
        _require_cuda(tiles, "count_violations")
        t = tiles if tiles.dim() == 3 else tiles.unsqueeze(0)
        t = t.to(torch.int64).contiguous()
        B, H, W = t.shape
        out = torch.empty((B, 2), dtype=torch.int64, device=t.device)
        nat.check(nat.lib().wfd_wfc_count_violations(self.handle, C.c_void_p(t.data_ptr()), B, H, W,
                                                     C.c_void_p(out.data_ptr()), nat.stream_ptr()),
                  "wfd_wfc_count_violations")
        return out

    def wave_view(self, B, H, W):
        ptr = nat.lib().wfd_wfc_saved_wave(self.handle, C.c_void_p(self._saved_ptr), B, H, W)
        off = ptr - self._saved.data_ptr()
        n = B * H * W * self.t_pad
        return self._saved[off:off + 2 * n].view(_torch16(self.dtype16)).view(B, H, W, self.t_pad)


def to_nhwc16(t, dtype16):
    """(B, C, H, W) float tensor of any layout -> contiguous [B, H, W, C] 16-bit tensor, via the library's kernels."""
    B, Cc, H, W = t.shape
    out = torch.empty((B, H, W, Cc), dtype=_torch16(dtype16), device=t.device)
    if t.dtype == _torch16(dtype16) and t.is_contiguous(memory_format=torch.channels_last):
        return t.permute(0, 2, 3, 1)
    src = t
    if src.dtype not in (torch.float32, torch.float16):
        src = src.float()
    if src.is_contiguous(memory_format=torch.channels_last) and not src.is_contiguous():
        # already channels-last in memory: a dtype conversion is all that is needed
        return src.permute(0, 2, 3, 1).to(_torch16(dtype16)).contiguous()
    src = src.contiguous()
    nat.check(nat.lib().wfd_nchw_to_nhwc16(C.c_void_p(src.data_ptr()), 0 if src.dtype == torch.float32 else 1,
                                           C.c_void_p(out.data_ptr()), B, Cc, H * W, dtype16, nat.stream_ptr()),
              "wfd_nchw_to_nhwc16")
    return out


def from_nhwc16(x, like, dtype16):
    """[B, H, W, C] 16-bit -> tensor shaped/typed like `like` ((B, C, H, W), fp32 or fp16)."""
    B, H, W, Cc = x.shape
    if like.dtype in (torch.float32, torch.float16) and like.is_contiguous():
        out = torch.empty((B, Cc, H, W), dtype=like.dtype, device=x.device)
        nat.check(nat.lib().wfd_nhwc16_to_nchw(C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()),
                                               0 if like.dtype == torch.float32 else 1, B, Cc, H * W, dtype16,
                                               nat.stream_ptr()), "wfd_nhwc16_to_nchw")
        return out
    return x.permute(0, 3, 1, 2).to(like.dtype)


class WfcLossFunction(torch.autograd.Function):
    """loss = WFCLossEinsum(wave) (reference wfd/wfdf/losses.py:63-75), wave given as a (B, T_pad, H, W) tensor."""

    @staticmethod
    def forward(ctx, t, handle, softmax):
        _require_cuda(t, "WFCLossEinsum")
        B, Cc, H, W = t.shape
        x = to_nhwc16(t, handle.dtype16)
        if not x.is_contiguous():
            x = x.contiguous()
        # the block backward needs belongs to THIS call (ctx keeps it alive): loss(a) + loss(b) differentiates both
        # against their own waves, and the handle's shared block (wave_view) is left to the fused guidance path
        saved, saved_ptr = handle.new_saved(B, H, W)
        loss = handle.forward(x.data_ptr(), softmax, B, H, W, 1.0, saved_ptr=saved_ptr)
        ctx.handle, ctx.x, ctx.like, ctx.saved, ctx.saved_ptr = handle, x, t, saved, saved_ptr
        return loss.to(t.dtype) if t.dtype != torch.float32 else loss

    @staticmethod
    def backward(ctx, dloss):
        h = ctx.handle
        x = ctx.x
        dout = torch.empty_like(x)
        h.backward(dloss.float().contiguous(), dout.data_ptr(), saved_ptr=ctx.saved_ptr)
        return from_nhwc16(dout, ctx.like, h.dtype16), None, None


def max_micro_batch():
    """Largest batch one library call carries (activation workspace ~0.6 GB per 64x64 map); larger batches are run
    as consecutive micro-batches (maps are independent). Override with WFD_MAX_BATCH."""
    return max(1, int(os.environ.get("WFD_MAX_BATCH", "64")))


def cuda_graphs_enabled():
    """The guidance step is a fixed launch sequence (no host sync, no allocation), so it is replayed from a CUDA graph
    once its shapes are known; WFD_CUDA_GRAPH=0 falls back to plain launches."""
    return os.environ.get("WFD_CUDA_GRAPH", "1") != "0"


class GraphedGuidanceStep:
    """One captured wfd_guidance_step for fixed (plan, handle, batch, dtype, scales): static input/output buffers."""

    def __init__(self, plan, handle, B, h, w, zdtype, in_scale, strength, device):
        self.z = torch.zeros((B, 4, h, w), dtype=zdtype, device=device)
        self.loss = torch.empty(B, dtype=torch.float32, device=device)
        self.dz = torch.empty((B, 4, h, w), dtype=torch.float32, device=device)
        import weakref

        self._plan, self.handle = weakref.ref(plan), handle  # the plan owns this object (plan._graphs): no cycle
        saved = handle.saved_ptr(B, plan.H, plan.W)
        zcode = _z_dtype_code(self.z)

        def run():
            nat.check(nat.lib().wfd_guidance_step(plan.handle, handle.handle, C.c_void_p(self.z.data_ptr()), zcode, B,
                                                  float(in_scale), float(strength), saved,
                                                  C.c_void_p(self.loss.data_ptr()), C.c_void_p(self.dz.data_ptr()),
                                                  nat.stream_ptr()), "wfd_guidance_step")

        cur = torch.cuda.current_stream(device)
        side = torch.cuda.Stream(device=device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            run()  # warm-up outside capture: function attributes, tensor maps
        cur.wait_stream(side)
        torch.cuda.synchronize(device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            run()
        self.saved_owner = handle._saved  # keep the saved block alive / detect reallocation

    def valid(self):
        return self.saved_owner is self.handle._saved

    def __call__(self, z):
        self.z.copy_(z)
        self._plan().generation += 1
        self.graph.replay()
        return self.loss, self.dz


class GraphedGuidanceLoop:
    """n_steps guided latent updates for an affine scheduler step (prev = a x + b eps) as ONE captured launch sequence:
    wfd_guidance_loop = the body of the reference's final-stage loop (pipeline_wfd.py:626-639) without a host round trip
    between its iterations. Static buffers; a, b, strength and n_steps are part of the capture."""

    def __init__(self, plan, handle, B, h, w, a, b, in_scale, strength, n_steps, device):
        import weakref

        self._plan, self.handle = weakref.ref(plan), handle
        self.lat = torch.zeros((B, 4, h, w), dtype=torch.float32, device=device)
        self.eps = torch.zeros_like(self.lat)
        self.z = torch.empty_like(self.lat)
        self.dz = torch.empty_like(self.lat)
        self.loss = torch.empty((n_steps, B), dtype=torch.float32, device=device)
        saved = handle.saved_ptr(B, plan.H, plan.W)

        def run(steps):
            nat.check(nat.lib().wfd_guidance_loop(plan.handle, handle.handle, C.c_void_p(self.lat.data_ptr()),
                                                  C.c_void_p(self.eps.data_ptr()), None, 0.0, float(a), float(b), B,
                                                  float(in_scale), float(strength), steps, saved,
                                                  C.c_void_p(self.loss.data_ptr()), C.c_void_p(self.z.data_ptr()),
                                                  C.c_void_p(self.dz.data_ptr()), nat.stream_ptr()), "wfd_guidance_loop")

        cur = torch.cuda.current_stream(device)
        side = torch.cuda.Stream(device=device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            run(1)  # warm-up outside capture: function attributes, tensor maps
        cur.wait_stream(side)
        torch.cuda.synchronize(device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            run(n_steps)
        self.saved_owner = handle._saved
        self.n_steps = n_steps

    def valid(self):
        return self.saved_owner is self.handle._saved

    def __call__(self, latents, eps):
        self.lat.copy_(latents)
        self.eps.copy_(eps)
        self._plan().generation += 1
        self.graph.replay()
        return self.lat.clone(), self.loss


def guidance_loop(latents, eps, native, handle, a, b, in_scale, strength, n_steps):
    """latents, eps: (B, 4, h, w) fp32 on the device -> (new latents, losses (n_steps, B)) after n_steps updates
    lat -= a * d/dz [strength * wfc_loss(softmax(decode((a lat + b eps) * in_scale)))]."""
    _require_cuda(latents, "guidance loop")
    B, _, h, w = latents.shape
    plan = native.plan(B, h, w, latents.device)
    cache = plan.__dict__.setdefault("_graphs", {})
    key = ("loop", id(handle), B, float(a), float(b), float(in_scale), float(strength), int(n_steps),
           torch.cuda.current_stream().cuda_stream)
    g = cache.get(key)
    if g is None or not g.valid() or g.handle is not handle:
        if len(cache) > 8:
            cache.clear()
        g = GraphedGuidanceLoop(plan, handle, B, h, w, a, b, in_scale, strength, n_steps, latents.device)
        cache[key] = g
    return g(latents.float(), eps.float())


class GuidanceLossFunction(torch.autograd.Function):
    """Fused guidance loss: loss[b] = strength * WFCLossEinsum(softmax(decode(latents * in_scale)))[b]
    (reference pipeline_wfd.py:612 with decode_latents_tile :402-408). One library call computes loss and d loss / d
    latents; backward just scales by the incoming gradient."""

    @staticmethod
    def forward(ctx, latents, native, handle, in_scale, strength):
        _require_cuda(latents, "guidance loss")
        B, _, h, w = latents.shape
        mb = min(B, max_micro_batch())
        plan = native.plan(mb, h, w, latents.device)
        z = latents.contiguous()
        if cuda_graphs_enabled() and B <= mb and not torch.cuda.is_current_stream_capturing():
            # captured steps live ON the plan (they die with it: NativeDecoder.invalidate() after .to()/.half()/
            # load_state_dict frees their workspace too); an entry keeps its handle alive, so id(handle) cannot be reused
            cache = plan.__dict__.setdefault("_graphs", {})
            key = (id(handle), B, z.dtype, float(in_scale), float(strength), torch.cuda.current_stream().cuda_stream)
            g = cache.get(key)
            if g is None or not g.valid() or g.handle is not handle:
                if len(cache) > 8:
                    cache.clear()
                g = GraphedGuidanceStep(plan, handle, B, h, w, z.dtype, in_scale, strength, z.device)
                cache[key] = g
            loss_s, dz_s = g(z)
            ctx.save_for_backward(dz_s.clone())
            ctx.zdtype = latents.dtype
            return loss_s.clone()
        loss = torch.empty(B, dtype=torch.float32, device=z.device)
        dz = torch.empty((B, 4, h, w), dtype=torch.float32, device=z.device)
        esz = z.element_size() * 4 * h * w
        for lo in range(0, B, mb):
            n = min(mb, B - lo)
            plan.generation += 1  # the plan's activation workspace is overwritten: an earlier decode cannot run backward
            nat.check(nat.lib().wfd_guidance_step(plan.handle, handle.handle, C.c_void_p(z.data_ptr() + lo * esz),
                                                  _z_dtype_code(z), n, float(in_scale), float(strength),
                                                  handle.saved_ptr(n, plan.H, plan.W), C.c_void_p(loss.data_ptr() + 4 * lo),
                                                  C.c_void_p(dz.data_ptr() + 16 * h * w * lo), nat.stream_ptr()),
                      "wfd_guidance_step")
        ctx.save_for_backward(dz)
        ctx.zdtype = latents.dtype
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (dz,) = ctx.saved_tensors
        g = dz * dloss.float().view(-1, 1, 1, 1)
        return g.to(ctx.zdtype), None, None, None, None


def argmax_tiles(x):
    """numpy.argmax(x, axis=-1) on the device: x [..., n] contiguous fp32/fp16/bf16 -> int64 [...]."""
    _require_cuda(x, "argmax_tiles")
    x = x.contiguous()
    n = x.shape[-1]
    rows = x.numel() // n if n > 0 else 0
    out = torch.empty(x.shape[:-1], dtype=torch.int64, device=x.device)
    if rows == 0:
        return out
    code = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}[x.dtype]
    nat.check(nat.lib().wfd_argmax_tiles(C.c_void_p(x.data_ptr()), code, rows, n, C.c_void_p(out.data_ptr()),
                                         nat.stream_ptr()), "wfd_argmax_tiles")
    return out
