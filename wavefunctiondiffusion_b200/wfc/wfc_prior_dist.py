"""WFCGeneratorPriorDist — the constraint state and information propagation of the reference's prior-distribution WFC
solver (wfd/wfc/wfc_prior_dist.py:3-43 constructor, :136-174 propagate_information, :176-205 double_check, :207-226 null
tile and border rules), with the allowed-tile sets held as bit sets on the GPU and propagated there for the WHOLE map at
once (libwfd_b200.so, csrc/propagate.cu). The reference author wanted to seed this solver with the diffusion wave
(README.md:149-151); `from_wave` does that seeding on the device.

What is mirrored: the constructor arguments and their meaning, `map_allowed_tiles`, `connections`, `boundary_grids`,
`is_grid_allowed / is_fixed / is_zero / get_possible_states`, `propagate_information` and `double_check` return values.
What differs, deliberately: propagation always runs to the fixed point over the whole map, every sweep from the previous
sweep's state (the reference's heap walk of bounded depth visits a neighbourhood in place; repeated until nothing changes
it ends in the same state as long as no cell runs empty, because the update is monotone), and a contradiction is
reported without rolling the state back. The tile choice / backtracking loop of the
solver (`generate*`, :360-664) stays on the host and is out of scope. There is no CPU path: without the CUDA library
every propagating call raises.
"""
import ctypes as C

import numpy as np
import torch

from .. import _native as nat

DIRECTIONS = ((-1, 0), (0, -1), (1, 0), (0, 1))  # (dx, dy): left, up, right, down - wfc_prior_dist.py:25


def connections_from_wfc_mats(wfc_mat_x, wfc_mat_y):
    """(T, 4, T) bool `connections` from the loss's adjacency matrices (x[i, j]: j allowed right of i; y[i, j]: j allowed
    below i - wfd/wfdf/losses.py:48-51): connections[s, d, t] = t may sit at direction d of s."""
    ax, ay = np.asarray(wfc_mat_x) != 0, np.asarray(wfc_mat_y) != 0
    T = ax.shape[0]
    c = np.zeros((T, 4, T), dtype=bool)
    c[:, 0, :] = ax.T  # t left of s   <=> s right of t
    c[:, 1, :] = ay.T  # t above s     <=> s below t
    c[:, 2, :] = ax    # t right of s
    c[:, 3, :] = ay    # t below s
    return c


def pack_bits(a):
    """bool (..., T) -> uint32 (..., ceil(T / 32)), bit (t & 31) of word (t >> 5)."""
    a = np.asarray(a, dtype=bool)
    T = a.shape[-1]
    Wd = (T + 31) // 32
    padded = np.zeros(a.shape[:-1] + (Wd * 32,), dtype=np.uint8)
    padded[..., :T] = a
    return np.packbits(padded, axis=-1, bitorder="little").view(np.uint32).reshape(a.shape[:-1] + (Wd,))


def unpack_bits(w, T):
    """uint32 (..., Wd) -> bool (..., T)."""
    w = np.ascontiguousarray(np.asarray(w, dtype=np.uint32))
    bits = np.unpackbits(w.view(np.uint8).reshape(w.shape[:-1] + (w.shape[-1] * 4,)), axis=-1, bitorder="little")
    return bits[..., :T].astype(bool)


def _u32_tensor(a, device):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).to(device)


class WFCGeneratorPriorDist:
    """
    Parameters (as wfd/wfc/wfc_prior_dist.py:4-8):
        map_size: (max_x, max_y).
        prior: (max_y, max_x, T) tile probabilities (the wave), kept for the host-side tile choice; may be None.
        connection_probs: (T, 4, T); > 0 means tile t may sit at direction d of tile s.
        boundary_grids: (max_y, max_x) bool, cells that take no part.
        preeliminate_border: remove tiles without any neighbour in a direction from cells that have one there.
    Extra: `batch` independent maps share the rules (the reference holds one map), `device`.
    """

    def __init__(self, map_size, prior, connection_probs=None, use_uniform_probs=False, boundary_grids=None,
                 preeliminate_border=True, max_saved_states=3, batch=1, device="cuda"):
        if connection_probs is None:
            raise ValueError("connection_probs is required")
        self.prior = prior
        self.map_size = map_size
        self.max_x, self.max_y = map_size
        self.batch = int(batch)
        self.device = torch.device(device)
        connections = np.array(np.asarray(connection_probs) > 0)
        self.tile_count = connections.shape[0]
        if connections.shape != (self.tile_count, 4, self.tile_count):
            raise ValueError("connection_probs must be (T, 4, T), got %s" % (connections.shape,))
        self.connections = connections
        self.connection_probs = connection_probs
        self.use_uniform_probs = use_uniform_probs
        self.directions = DIRECTIONS
        self.max_saved_states = max_saved_states
        self.info_prop_count = 0
        self.last_sweeps = 0
        allowed = np.ones((self.max_y, self.max_x, self.tile_count), dtype=bool)
        if preeliminate_border:
            self._preeliminate_border_tiles(allowed)
        # allow_all_connections_to_null (:207-212): the null tile connects to everything and is never placed
        self.connections[0, :, :] = True
        self.connections[:, :, 0] = True
        allowed[:, :, 0] = False
        self.boundary_grids = (np.zeros((self.max_y, self.max_x), dtype=bool) if boundary_grids is None
                               else np.asarray(boundary_grids, dtype=bool))
        self._handle = None
        self._boundary_dev = None
        self._allowed_bits = None
        self.set_allowed(np.broadcast_to(allowed, (self.batch,) + allowed.shape))

    # ------------------------------------------------------------------ host-side rules (plain numpy, no propagation)
    def _preeliminate_border_tiles(self, allowed):
        """:214-224 - a tile with no allowed neighbour at all on one side can only sit on that border."""
        c = self.connections
        allowed[:, 1:, np.where(c[:, 0, :].sum(axis=1) == 0)[0]] = False
        allowed[1:, :, np.where(c[:, 1, :].sum(axis=1) == 0)[0]] = False
        allowed[:, :-1, np.where(c[:, 2, :].sum(axis=1) == 0)[0]] = False
        allowed[:-1, :, np.where(c[:, 3, :].sum(axis=1) == 0)[0]] = False

    def is_grid_allowed(self, grid):
        y, x = grid
        return 0 <= x < self.max_x and 0 <= y < self.max_y and not self.boundary_grids[grid]

    # ------------------------------------------------------------------ device state
    def _require_lib(self, what):
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise nat.WfdError("%s runs on the GPU only (device %s); there is no CPU path" % (what, self.device))
        return nat.lib()

    def _rules(self):
        """The native propagator handle (connection rules re-encoded for the device) and the boundary mask, built once."""
        if self._handle is None:
            lib = nat.lib()
            bits = np.ascontiguousarray(pack_bits(np.ascontiguousarray(self.connections.transpose(1, 0, 2))))  # [4, T, Wd], host
            h = C.c_void_p()
            with torch.cuda.device(self.device):
                nat.check(lib.wfd_propagator_create(bits.ctypes.data_as(C.c_void_p), self.tile_count, C.byref(h)),
                          "wfd_propagator_create")
            self._handle = h
            self._boundary_dev = torch.from_numpy(np.ascontiguousarray(
                np.broadcast_to(self.boundary_grids, (self.batch,) + self.boundary_grids.shape)).astype(np.uint8)).to(self.device)
        return self._handle

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                nat.lib().wfd_propagator_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    def set_allowed(self, allowed):
        """allowed: bool (max_y, max_x, T) or (batch, max_y, max_x, T); replaces the state."""
        a = np.asarray(allowed, dtype=bool)
        if a.ndim == 3:
            a = np.broadcast_to(a, (self.batch,) + a.shape)
        if a.shape != (self.batch, self.max_y, self.max_x, self.tile_count):
            raise ValueError("allowed must be (batch, max_y, max_x, T) = %s, got %s"
                             % ((self.batch, self.max_y, self.max_x, self.tile_count), a.shape))
        bits = pack_bits(a)
        self._allowed_bits = _u32_tensor(bits, self.device) if self.device.type == "cuda" and torch.cuda.is_available() \
            else torch.from_numpy(np.ascontiguousarray(bits).view(np.int32))

    def set_allowed_bits(self, bits):
        """bits: int32 tensor (batch, max_y, max_x, Wd) already on the device (e.g. from `wave_to_allowed`)."""
        Wd = (self.tile_count + 31) // 32
        if tuple(bits.shape) != (self.batch, self.max_y, self.max_x, Wd) or bits.dtype != torch.int32:
            raise ValueError("bits must be int32 (batch, max_y, max_x, %d)" % Wd)
        self._allowed_bits = bits.contiguous()

    @property
    def allowed_bits(self):
        return self._allowed_bits

    @property
    def map_allowed_tiles(self):
        """bool (max_y, max_x, T) for one map, (batch, max_y, max_x, T) for several - a host copy of the device state."""
        a = unpack_bits(self._allowed_bits.cpu().numpy().view(np.uint32), self.tile_count)
        return a[0] if self.batch == 1 else a

    @classmethod
    def from_wave(cls, wave, wfc_mat_x, wfc_mat_y, threshold=1e-3, keep_argmax=True, boundary_grids=None,
                  preeliminate_border=True):
        """wave: (B, H, W, C >= T) probabilities on the device, channels last (what `decode_latents_tile` produces before
        it leaves the GPU; fp32 / fp16 / bf16). A tile stays possible in a cell when its probability reaches `threshold`
        (the cell's most probable tile always stays with keep_argmax), intersected with the constructor's rules."""
        if not wave.is_cuda:
            raise nat.WfdError("from_wave needs a CUDA tensor; there is no CPU path")
        if wave.dim() == 3:
            wave = wave.unsqueeze(0)
        B, H, W, Cw = wave.shape
        conn = connections_from_wfc_mats(wfc_mat_x, wfc_mat_y)
        gen = cls((W, H), None, conn.astype(np.float32), boundary_grids=boundary_grids,
                  preeliminate_border=preeliminate_border, batch=B, device=wave.device)
        gen.prior = wave
        seeded = wave_to_allowed(wave, gen.tile_count, threshold, keep_argmax, first_tile=1)
        gen.set_allowed_bits(torch.bitwise_and(seeded, gen._allowed_bits))
        return gen

    # ------------------------------------------------------------------ propagation
    def propagate_all(self, max_sweeps=1024):
        """Propagates every cell's information to the fixed point. Returns False when a cell is left without any tile
        (what propagate_information / double_check report, :152-153, :166-167), True otherwise."""
        lib = self._require_lib("WFCGeneratorPriorDist.propagate_all")
        handle = self._rules()
        B, H, W = self.batch, self.max_y, self.max_x
        sweeps, conv = C.c_int(0), C.c_int(0)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            nat.check(lib.wfd_propagator_run(handle, C.c_void_p(self._allowed_bits.data_ptr()),
                                             C.c_void_p(self._boundary_dev.data_ptr()), B, H, W, int(max_sweeps),
                                             C.byref(sweeps), C.byref(conv), C.c_void_p(stream)), "wfd_propagator_run")
        self.last_sweeps, self.converged = sweeps.value, bool(conv.value)
        self.info_prop_count += sweeps.value * B * H * W
        counts = self.count_allowed()
        live = ~self._boundary_dev.bool()
        return not bool(((counts == 0) & live).any())

    def propagate_information(self, main_grid=None, depth=None, borders=None, updated_tiles=None, verbose=False):
        """Reference signature (:136). The device propagation covers the whole map to the fixed point, which contains
        whatever a bounded-depth walk from `main_grid` would have removed; the arguments are accepted and ignored."""
        return self.propagate_all()

    def double_check(self, check_rect=None, verbose=False):
        """:176-205 - propagate from every cell; False on a contradiction, else whether every cell is decided."""
        if not self.propagate_all():
            return False
        counts = self.count_allowed()
        self._rules()
        live = ~self._boundary_dev.bool()
        if check_rect is not None:
            x1, y1, x2, y2 = check_rect
            sel = torch.zeros_like(live)
            sel[:, y1:y2, x1:x2] = True
            live = live & sel
        return bool(((counts == 1) | ~live).all())

    def count_allowed(self):
        """int32 (batch, max_y, max_x) on the device: tiles each cell still allows."""
        lib = self._require_lib("WFCGeneratorPriorDist.count_allowed")
        B, H, W, T = self.batch, self.max_y, self.max_x, self.tile_count
        out = torch.empty((B, H, W), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            nat.check(lib.wfd_wfc_allowed_count(C.c_void_p(self._allowed_bits.data_ptr()), B, H, W, T,
                                                C.c_void_p(out.data_ptr()), C.c_void_p(stream)), "wfd_wfc_allowed_count")
        return out

    def _cell_bits(self, grid, b=0):
        y, x = grid
        return unpack_bits(self._allowed_bits[b, y, x].cpu().numpy().view(np.uint32), self.tile_count)

    def is_fixed(self, grid, b=0):
        return int(self._cell_bits(grid, b).sum()) == 1

    def is_zero(self, grid, b=0):
        return int(self._cell_bits(grid, b).sum()) == 0

    def get_possible_states(self, grid, b=0):
        return np.where(self._cell_bits(grid, b))[0]

    def get_first_possible_state(self, grid, b=0):
        return int(self._cell_bits(grid, b).argmax())


def wave_to_allowed(wave, tile_count, threshold, keep_argmax=True, first_tile=1):
    """wave (B, H, W, C >= T) on the device -> int32 bit sets (B, H, W, ceil(T / 32)); see wfd_wave_to_allowed."""
    if not wave.is_cuda:
        raise nat.WfdError("wave_to_allowed needs a CUDA tensor; there is no CPU path")
    dt = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}.get(wave.dtype)
    if dt is None:
        raise ValueError("wave dtype %s not supported" % wave.dtype)
    wave = wave.contiguous()
    B, H, W, Cw = wave.shape
    if Cw < tile_count:
        raise ValueError("wave has %d channels, fewer than the %d tiles" % (Cw, tile_count))
    Wd = (tile_count + 31) // 32
    out = torch.empty((B, H, W, Wd), dtype=torch.int32, device=wave.device)
    with torch.cuda.device(wave.device):
        stream = torch.cuda.current_stream().cuda_stream
        nat.check(nat.lib().wfd_wave_to_allowed(C.c_void_p(wave.data_ptr()), dt, B * H * W, Cw, tile_count, int(first_tile),
                                                float(threshold), int(bool(keep_argmax)), C.c_void_p(out.data_ptr()),
                                                C.c_void_p(stream)), "wfd_wave_to_allowed")
    return out
