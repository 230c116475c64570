"""TEST INFRASTRUCTURE ONLY - CPU restatement (numpy) of the information propagation of the reference's prior-distribution
WFC solver, the checker for csrc/propagate.cu. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
import this file; the product (wavefunctiondiffusion_b200/wfc) never does.

Follows wfd/wfc/wfc_prior_dist.py:
  :3-43     constructor: connections = connection_probs > 0, border pre-elimination (:214-224) BEFORE the null-tile rule
            (:207-212: tile 0 connects to everything and is never placed)
  :136-174  propagate_information: for a source cell with states S and every direction i with an allowed target cell,
            target &= OR_{s in S} connections[s, i, :]; an empty source returns False without touching anything
  :176-205  double_check: that step (depth 1) from every allowed cell in raster order
Pinned by tests/golden/wfc_propagate.npz, which oracle/make_golden.py writes by running the UNMODIFIED reference class
(imported from /root/reference) until its state stops changing: tests/test_oracle.py compares this file with it.

The closure here is computed Jacobi-style (all cells from the previous state at once). The update is monotone, so the
fixed point it reaches is the one the reference's raster order reaches.
"""
import numpy as np

DIRECTIONS = ((-1, 0), (0, -1), (1, 0), (0, 1))  # (dx, dy), wfc_prior_dist.py:25


def connections_from_wfc_mats(mat_x, mat_y):
    """connections[s, d, t] from x[i, j] (j allowed right of i) and y[i, j] (j allowed below i), losses.py:48-51."""
    ax, ay = np.asarray(mat_x) != 0, np.asarray(mat_y) != 0
    return np.stack([ax.T, ay.T, ax, ay], axis=1)


def constructor_state(map_size, connection_probs, preeliminate_border=True):
    """(connections, map_allowed_tiles) as the reference constructor leaves them (:11-33)."""
    max_x, max_y = map_size
    c = np.array(np.asarray(connection_probs) > 0)
    T = c.shape[0]
    allowed = np.ones((max_y, max_x, T), dtype=bool)
    if preeliminate_border:
        no_nb = [np.where(c[:, d, :].sum(axis=1) == 0)[0] for d in range(4)]
        allowed[:, 1:, no_nb[0]] = False
        allowed[1:, :, no_nb[1]] = False
        allowed[:, :-1, no_nb[2]] = False
        allowed[:-1, :, no_nb[3]] = False
    c[0, :, :] = True
    c[:, :, 0] = True
    allowed[:, :, 0] = False
    return c, allowed


def propagate_sweep(allowed, connections, boundary=None):
    """One Jacobi sweep over a (H, W, T) bool state: every cell narrowed by its four in-map sources' previous states."""
    H, W, T = allowed.shape
    live = np.ones((H, W), dtype=bool) if boundary is None else ~np.asarray(boundary, dtype=bool)
    out = allowed.copy()
    flat = allowed.reshape(H * W, T).astype(np.float32)
    for d, (dx, dy) in enumerate(DIRECTIONS):
        support = (flat @ connections[:, d, :].astype(np.float32) > 0).reshape(H, W, T)
        nonempty = allowed.any(axis=2) & live
        for y in range(H):
            for x in range(W):
                ny, nx = y + dy, x + dx
                if not nonempty[y, x] or not (0 <= ny < H and 0 <= nx < W) or not live[ny, nx]:
                    continue
                out[ny, nx] &= support[y, x]
    return out


def propagate_closure(allowed, connections, boundary=None, max_sweeps=100000):
    """Fixed point of propagate_sweep. Returns (state, ok, sweeps); ok is False when a live cell is left empty."""
    cur = np.asarray(allowed, dtype=bool).copy()
    sweeps = 0
    while sweeps < max_sweeps:
        nxt = propagate_sweep(cur, connections, boundary)
        sweeps += 1
        if (nxt == cur).all():
            break
        cur = nxt
    live = np.ones(cur.shape[:2], dtype=bool) if boundary is None else ~np.asarray(boundary, dtype=bool)
    return cur, not bool((~cur.any(axis=2) & live).any()), sweeps


def wave_to_allowed(wave, T, threshold, keep_argmax=True, first_tile=1):
    """allowed[..., t] = first_tile <= t < T and (wave[..., t] >= threshold or t == argmax over [first_tile, T))."""
    w = np.asarray(wave, dtype=np.float32)[..., :T]
    out = w >= np.float32(threshold)
    if keep_argmax:
        am = w[..., first_tile:].argmax(axis=-1) + first_tile
        np.put_along_axis(out, am[..., None], True, axis=-1)
    out[..., :first_tile] = False
    return out
