"""WFCLossEinsum — host-side mirror of the reference loss module (wfd/wfdf/losses.py:43-75).

Same constructor, buffers (`cx`, `cy` = 1 - adjacency, shaped (1, T, T, 1, 1)), `count` attribute and call signature
`loss(t, softmax=False) -> (B,)`; the contraction itself runs in libwfd_b200.so (tcgen05 kernel against the four
directional adjacency matrices) and is differentiable w.r.t. `t` through a torch.autograd.Function.
"""
import torch

from ..wf_diffusers.native_ops import WfcHandle, WfcLossFunction, _require_cuda, compute_dtype16


class WFCLossEinsum(torch.nn.Module):
    """
    Expectation of violating the WFC transition rules on a wave.

    Parameters:
        wfc_mat_x: (T, T), x[i, j] == 1 means tile j is allowed at the right side of tile i.
        wfc_mat_y: (T, T), y[i, j] == 1 means tile j is allowed below tile i.
    """

    def __init__(self, wfc_mat_x, wfc_mat_y):
        super().__init__()
        cx = torch.as_tensor(wfc_mat_x, dtype=torch.float32)
        cy = torch.as_tensor(wfc_mat_y, dtype=torch.float32)
        assert cx.size(0) == cx.size(1) and cx.size(1) == cy.size(0) and cy.size(0) == cy.size(1), "WFC matrix size mismatch"
        self.register_buffer("cx", (1 - cx)[None, :, :, None, None])
        self.register_buffer("cy", (1 - cy)[None, :, :, None, None])
        self.count = cx.size(0)
        self._handles = {}
        self.sparse_contraction = False  # set_sparse(): the CUDA-core gather over allowed pairs instead of the dense product

    # copies and pickles carry the buffers only; native handles are rebuilt on first use
    def __getstate__(self):
        state = self.__dict__.copy()
        state["_handles"] = {}
        return state

    def __deepcopy__(self, memo):
        import copy

        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        new.__dict__.update({k: copy.deepcopy(v, memo) for k, v in self.__getstate__().items()})
        return new

    def _mats(self):
        ax = (self.cx[0, :, :, 0, 0] < 0.5).to("cpu").numpy()
        ay = (self.cy[0, :, :, 0, 0] < 0.5).to("cpu").numpy()
        return ax, ay

    def native_handle(self, t_pad, device):
        key = (t_pad, compute_dtype16(), str(device))
        h = self._handles.get(key)
        if h is None:
            ax, ay = self._mats()
            h = WfcHandle(ax, ay, t_pad, key[1], device)
            if self.sparse_contraction:
                h.set_sparse(True)
            self._handles[key] = h
        return h

    def set_sparse(self, on=True):
        """Not part of the reference API: score through the sparse complement form
        p^T (1 - A) q = sum(p) sum(q) - p^T A q with A as lists of allowed pairs (a CUDA-core gather) instead of the dense
        tensor-core contraction. Loss and gradient agree with the dense path to rounding."""
        self.sparse_contraction = bool(on)
        for h in self._handles.values():
            h.set_sparse(self.sparse_contraction)
        return self

    def count_violations(self, tiles):
        """Forbidden adjacent pairs of integer tile maps (B, H, W) -> int64 (B, 2) = (horizontal, vertical): the
        one-hot case of forward(), counted on the device (not part of the reference API)."""
        _require_cuda(tiles, "WFCLossEinsum.count_violations")
        t_pad = (self.count + 127) // 128 * 128
        return self.native_handle(t_pad, tiles.device).count_violations(tiles)

    def forward(self, t, softmax=False):
        if t.dim() != 4 or t.size(1) < self.count:
            raise ValueError(f"wave must be (B, C>={self.count}, H, W), got {tuple(t.shape)}")
        _require_cuda(t, "WFCLossEinsum")
        # The reference accepts any channel count >= count (losses.py:63-66 slices [:count]); the contraction here walks the
        # channels in blocks of 64, so other counts (e.g. a wave with exactly `count` channels) are padded with channels
        # that carry no probability: zeros for a wave, a large negative logit when the softmax is still to come. The pad
        # is an ordinary differentiable torch op, its backward is the slice.
        pad = (-t.size(1)) % 64
        if pad:
            t = torch.nn.functional.pad(t, (0, 0, 0, 0, 0, pad), value=(-3.0e4 if softmax else 0.0))
        handle = self.native_handle(t.size(1), t.device)
        return WfcLossFunction.apply(t, handle, bool(softmax))
