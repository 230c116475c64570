"""GPU: bit-set constraint propagation (csrc/propagate.cu through the C ABI) against the golden states of the unmodified
reference solver, against the numpy oracle on seeded cases, and through size-independent properties at the full ice size
(idempotence, monotonicity, arc consistency of sampled cells). Everything here is bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import wfc_propagate_oracle as porc
from wavefunctiondiffusion_b200.utils.tilesets import load_wfc_mats
from wavefunctiondiffusion_b200.wfc import WFCGeneratorPriorDist, connections_from_wfc_mats, pack_bits, unpack_bits
from wavefunctiondiffusion_b200.wfc.wfc_prior_dist import wave_to_allowed

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "wfc_propagate.npz"))


def test_golden_synthetic_with_boundary_cells(gold):
    gen = WFCGeneratorPriorDist((12, 10), None, gold["a_conn_probs"], boundary_grids=gold["a_boundary"])
    assert (gen.map_allowed_tiles == gold["a_allowed_ctor"]).all()
    gen.set_allowed(gold["a_allowed_in"])
    assert gen.propagate_all() is True
    assert gen.converged and gen.last_sweeps >= 3
    assert (gen.map_allowed_tiles == gold["a_allowed_out"]).all()
    # the reference's double_check: no contradiction, but the map is not decided yet
    assert gen.double_check() is False
    assert gen.last_sweeps == 1  # already at the fixed point: one sweep finds nothing to change


def test_golden_ice_rules(gold):
    mx, my = load_wfc_mats("ice_64x64")[:2]
    gen = WFCGeneratorPriorDist((6, 6), None, connections_from_wfc_mats(mx, my).astype(np.float32))
    T = gen.tile_count
    gen.set_allowed(unpack_bits(gold["b_allowed_in"], T))
    assert gen.propagate_all() is True
    got = gen.allowed_bits.cpu().numpy().view(np.uint32)[0]
    assert (got == gold["b_allowed_out"]).all()
    assert gen.is_fixed((2, 3)) and not gen.is_zero((0, 0))
    counts = gen.count_allowed().cpu().numpy()[0]
    assert (counts == unpack_bits(gold["b_allowed_out"], T).sum(axis=2)).all()


def test_golden_contradiction(gold):
    gen = WFCGeneratorPriorDist((5, 4), None, gold["a_conn_probs"])
    gen.set_allowed(gold["c_allowed_in"])
    assert gen.propagate_all() is False and not bool(gold["c_ok"])
    assert gen.double_check() is False
    # same final state as the oracle, which also runs on to the fixed point instead of stopping at the first empty cell
    out, ok, _ = porc.propagate_closure(gold["c_allowed_in"], gold["a_connections"])
    assert not ok and (gen.map_allowed_tiles == out).all()


@pytest.mark.parametrize("T,H,W,B,p", [(70, 9, 7, 3, 0.4), (200, 1, 17, 2, 0.3), (33, 16, 1, 1, 0.5), (130, 2, 2, 2, 0.6)])
def test_against_the_oracle_seeded(T, H, W, B, p):
    rng = np.random.RandomState(T + H)
    src = rng.randint(1, T, size=(3 * H + 8, 3 * W + 8))
    probe = np.zeros((T, 4, T), dtype=np.float32)
    for d, (dx, dy) in enumerate(porc.DIRECTIONS):  # rules of a random map, as get_probs_from_map counts them
        a = src[max(0, -dy):src.shape[0] - max(0, dy), max(0, -dx):src.shape[1] - max(0, dx)]
        b = src[max(0, dy):src.shape[0] + min(0, dy), max(0, dx):src.shape[1] + min(0, dx)]
        probe[a.ravel(), d, b.ravel()] = 1
    boundary = rng.rand(H, W) < 0.08
    gen = WFCGeneratorPriorDist((W, H), None, probe, boundary_grids=boundary, batch=B)
    conn, ctor = porc.constructor_state((W, H), probe)
    assert (gen.connections == conn).all()
    seeds = np.stack([(rng.rand(H, W, T) < p) & ctor for _ in range(B)])
    for b in range(B):
        win = src[b:b + H, 2 * b:2 * b + W]
        np.put_along_axis(seeds[b], win[..., None], True, axis=2)
        seeds[b] &= ctor | np.eye(T, dtype=bool)[win]
    gen.set_allowed(seeds)
    ok = gen.propagate_all()
    got = gen.map_allowed_tiles.reshape(B, H, W, T)
    oks = []
    for b in range(B):
        out, ok_b, _ = porc.propagate_closure(seeds[b], conn, boundary)
        assert (got[b] == out).all(), "map %d differs" % b
        oks.append(ok_b)
    assert ok == all(oks)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_wave_to_allowed_against_the_oracle(dtype):
    g = torch.Generator().manual_seed(3)
    T, C = 3074, 3200
    wave = torch.softmax(4.0 * torch.randn(2, 5, 7, C, generator=g), dim=-1).to(dtype)
    wave[0, 0, 0, :] = 0  # all equal: the first tile of [first_tile, T) is the argmax
    wave[0, 0, 1, 5] = wave[0, 0, 1, 9] = 1.0  # a tie: the first one wins
    thr = 2e-3
    got = wave_to_allowed(wave.cuda(), T, thr, keep_argmax=True, first_tile=1).cpu().numpy().view(np.uint32)
    ref = porc.wave_to_allowed(wave.float().numpy(), T, thr, keep_argmax=True, first_tile=1)
    assert (got == pack_bits(ref)).all()
    assert unpack_bits(got, T)[0, 0, 0].tolist() == [False, True] + [False] * (T - 2)
    got = wave_to_allowed(wave.cuda(), T, thr, keep_argmax=False, first_tile=0).cpu().numpy().view(np.uint32)
    assert (got == pack_bits(porc.wave_to_allowed(wave.float().numpy(), T, thr, keep_argmax=False, first_tile=0))).all()


def test_full_size_ice_properties():
    """BASELINE size (ice, 64 x 64, batch 8) seeded from a wave: properties that hold at any size."""
    mx, my = load_wfc_mats("ice_64x64")[:2]
    T, C, B, H, W = mx.shape[0], 3200, 8, 64, 64
    g = torch.Generator(device="cuda").manual_seed(5)
    wave = torch.softmax(6.0 * torch.randn(B, H, W, C, generator=g, device="cuda"), dim=-1).to(torch.bfloat16)
    gen = WFCGeneratorPriorDist.from_wave(wave, mx, my, threshold=2e-3, keep_argmax=True)
    before = gen.allowed_bits.clone()
    ok = gen.propagate_all(max_sweeps=2048)
    assert gen.converged
    after = gen.allowed_bits.clone()
    assert bool((torch.bitwise_and(after, before) == after).all())  # monotone: only removals
    assert gen.propagate_all() == ok and gen.last_sweeps == 1         # idempotent
    assert bool((gen.allowed_bits == after).all())
    # arc consistency of sampled cells, recomputed on the host from the device state
    conn = gen.connections
    state = unpack_bits(after.cpu().numpy().view(np.uint32), T)
    rng = np.random.RandomState(1)
    for _ in range(24):
        b, y, x = rng.randint(B), rng.randint(H), rng.randint(W)
        for d, (dx, dy) in enumerate(porc.DIRECTIONS):
            sy, sx = y - dy, x - dx
            if not (0 <= sy < H and 0 <= sx < W) or not state[b, sy, sx].any():
                continue
            support = conn[state[b, sy, sx], d].any(axis=0)
            assert not (state[b, y, x] & ~support).any()
    counts = gen.count_allowed()
    assert bool((counts.cpu() == torch.from_numpy(state.sum(axis=-1).astype(np.int32))).all())
    assert ok == bool((counts > 0).all())
