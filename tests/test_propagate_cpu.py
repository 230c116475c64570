"""CPU: the propagation oracle (oracle/wfc_propagate_oracle.py) against the golden states written by the UNMODIFIED
reference solver class (oracle/make_golden.py propagate_fixture), and the host-side logic of the product mirror
(constructor rules, bit packing, loud failure without a GPU). No device calls."""
import os

import numpy as np
import pytest
import torch

from oracle import wfc_propagate_oracle as porc
from wavefunctiondiffusion_b200 import _native as nat
from wavefunctiondiffusion_b200.utils.tilesets import load_wfc_mats
from wavefunctiondiffusion_b200.wfc import WFCGeneratorPriorDist, connections_from_wfc_mats, pack_bits, unpack_bits

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "wfc_propagate.npz"))


def test_oracle_constructor_state_matches_the_reference(gold):
    conn, allowed = porc.constructor_state((12, 10), gold["a_conn_probs"])
    assert (conn == gold["a_connections"]).all()
    assert (allowed == gold["a_allowed_ctor"]).all()


def test_oracle_closure_matches_the_reference_synthetic(gold):
    out, ok, sweeps = porc.propagate_closure(gold["a_allowed_in"], gold["a_connections"], gold["a_boundary"])
    assert ok and bool(gold["a_ok"])
    assert (out == gold["a_allowed_out"]).all()          # bit-exact
    assert sweeps > 2 and out.sum() < gold["a_allowed_in"].sum()
    # boundary cells are neither narrowed nor used as sources
    b = gold["a_boundary"]
    assert (out[b] == gold["a_allowed_in"][b]).all()


@pytest.mark.slow
def test_oracle_closure_matches_the_reference_ice(gold):
    mx, my = load_wfc_mats("ice_64x64")[:2]
    conn, _ = porc.constructor_state((6, 6), porc.connections_from_wfc_mats(mx, my))
    T = conn.shape[0]
    out, ok, _ = porc.propagate_closure(unpack_bits(gold["b_allowed_in"], T), conn)
    assert ok and bool(gold["b_ok"])
    assert (pack_bits(out) == gold["b_allowed_out"]).all()


def test_oracle_reports_the_contradiction(gold):
    _, ok, _ = porc.propagate_closure(gold["c_allowed_in"], gold["a_connections"])
    assert not ok and not bool(gold["c_ok"])


def test_oracle_connections_orientation():
    mx = np.zeros((3, 3), dtype=np.uint8)
    my = np.zeros((3, 3), dtype=np.uint8)
    mx[1, 2] = 1  # 2 may sit right of 1
    my[2, 1] = 1  # 1 may sit below 2
    for fn in (porc.connections_from_wfc_mats, connections_from_wfc_mats):
        c = fn(mx, my)
        assert c.shape == (3, 4, 3)
        assert c[1, 2, 2] and c[2, 0, 1] and c[2, 3, 1] and c[1, 1, 2]
        assert c.sum() == 4


def test_bit_packing_round_trip():
    rng = np.random.RandomState(0)
    for T in (1, 31, 32, 33, 97, 3074):
        a = rng.rand(3, 2, T) < 0.3
        w = pack_bits(a)
        assert w.dtype == np.uint32 and w.shape == (3, 2, (T + 31) // 32)
        assert (unpack_bits(w, T) == a).all()
        t = T - 1
        one = np.zeros((T,), dtype=bool)
        one[t] = True
        assert pack_bits(one)[t >> 5] == np.uint32(1) << np.uint32(t & 31)


def test_product_constructor_matches_the_reference(gold):
    gen = WFCGeneratorPriorDist((12, 10), None, gold["a_conn_probs"], boundary_grids=gold["a_boundary"], device="cpu")
    assert (gen.connections == gold["a_connections"]).all()
    assert (gen.map_allowed_tiles == gold["a_allowed_ctor"]).all()
    assert gen.tile_count == 48 and (gen.max_x, gen.max_y) == (12, 10)
    assert not gen.is_grid_allowed((3, 4)) and gen.is_grid_allowed((0, 0)) and not gen.is_grid_allowed((10, 0))
    gen.set_allowed(gold["a_allowed_in"])
    assert (gen.map_allowed_tiles == gold["a_allowed_in"]).all()
    assert list(gen.get_possible_states((0, 0))) == list(np.where(gold["a_allowed_in"][0, 0])[0])


def test_propagation_fails_loudly_without_a_gpu(gold):
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    gen = WFCGeneratorPriorDist((12, 10), None, gold["a_conn_probs"], device="cpu")
    with pytest.raises(nat.WfdError):
        gen.propagate_all()
    with pytest.raises(nat.WfdError):
        WFCGeneratorPriorDist.from_wave(torch.zeros(1, 2, 2, 48), np.ones((48, 48)), np.ones((48, 48)))


def test_oracle_wave_to_allowed_rules():
    w = np.array([[0.5, 0.2, 0.2, 0.1], [0.9, 0.04, 0.04, 0.02], [0.0, 0.3, 0.3, 0.4]], dtype=np.float32)
    a = porc.wave_to_allowed(w, 4, 0.25, keep_argmax=True, first_tile=1)
    assert a.tolist() == [[False, True, False, False],      # below threshold everywhere: the first maximum of [1, T) stays
                          [False, True, False, False],
                          [False, True, True, True]]
    a = porc.wave_to_allowed(w, 3, 0.25, keep_argmax=False, first_tile=0)
    assert a.tolist() == [[True, False, False], [True, False, False], [False, True, True]]
