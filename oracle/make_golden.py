"""TEST INFRASTRUCTURE ONLY — generates tests/golden/* by running the UNMODIFIED reference modules
(/root/reference, imported through oracle/reference_loader.py) on seeded inputs. Run in the authoring container:

    python oracle/make_golden.py [--skip-full]

Fixtures (all small; weights are NOT stored — they are re-created from torch.manual_seed(seed) by the product's
AutoencoderTile, whose initialisation order is checked to be bit-identical to the reference's in this script):
  (package) data/wfc_mats/<tileset>.npz   sparse coordinates of the shipped wfc_mats (the adjacency input data of the path) + KATs
  tiny_step.npz            reference decode+softmax+loss+grad on a small config, fp64 and fp32
  tiny_down_step.npz       same with a DownDecoderBlock2D (the 32x32-style config)
  ice_c1_step.npz          BASELINE config 1: ice 64x64, batch 1, tilenet_sc_space_64x64, strength 5
  tiny_encode.npz          reference AutoencoderTile.encode moments of a one-hot tile map and of a soft wave (tiny config)
  tiles_post.npz           the reference's add_symmetry + gen_to_sc[shrink_to_gen[.]] on random ice tiles (with the tables)
  ice_c2_step.npz          BASELINE config 2's benchmarked shape: the same weights, batch 8 (slot 0 = the C1 latents)
  wfc_propagate.npz        the reference's WFCGeneratorPriorDist (wfd/wfc/wfc_prior_dist.py) propagated until its state stops
                           changing: a synthetic tileset learnt from a random map, the ice rules, a contradiction
"""
import glob
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.reference_loader import REFERENCE_ROOT, load_reference  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

TINY = dict(
    down_block_types=["EvenEncoderBlock2D"] * 4, up_block_types=["EvenDecoderBlock2D"] * 4,
    block_out_channels=[512, 256, 128, 64], conv_groups=[8, 4, 2, 1], conv_init_groups=16, layers_per_block=2,
    act_fn="silu", latent_channels=4, norm_num_groups=8, sample_size=16, in_channels=128, out_channels=128,
)
TINY_DOWN = dict(TINY, down_block_types=["EvenEncoderBlock2D", "UpEncoderBlock2D", "EvenEncoderBlock2D", "EvenEncoderBlock2D"],
                 up_block_types=["EvenDecoderBlock2D", "EvenDecoderBlock2D", "DownDecoderBlock2D", "EvenDecoderBlock2D"])


def synthetic_mats(T, density, seed):
    rng = np.random.RandomState(seed)
    m = rng.rand(2, T, T) < density
    m[:, 0, :] = False  # null tile: everything forbidden, like the shipped matrices
    m[:, :, 0] = False
    return m


def reference_step(vae, losses, cfg, mats, latents, strength, dtype, seed):
    torch.manual_seed(seed)
    net = vae.AutoencoderTile(**cfg).eval().to(dtype)
    lossm = losses.WFCLossEinsum(mats[0], mats[1]).to(dtype)
    lat = latents.to(dtype).requires_grad_()
    t0 = time.time()
    logits = net.decode(1 / 0.18215 * lat).sample      # pipeline_wfd.py:403-404
    wave = logits.softmax(axis=1)                       # :405
    loss = lossm(wave) * strength                       # :612
    (grad,) = torch.autograd.grad(loss.sum(), lat)      # :613 (.sum(): batch > 1 needs a scalar)
    dt = time.time() - t0
    return net, logits.detach(), wave.detach(), loss.detach(), grad.detach(), dt


def check_seed_parity(vae, cfg, seed):
    from wavefunctiondiffusion_b200.wf_diffusers.wf_vae import AutoencoderTile

    torch.manual_seed(seed)
    ref = vae.AutoencoderTile(**cfg)
    torch.manual_seed(seed)
    mine = AutoencoderTile(**cfg)
    a, b = ref.state_dict(), mine.state_dict()
    assert list(a.keys()) == list(b.keys()), "state_dict keys/order differ from the reference"
    for k in a:
        assert torch.equal(a[k], b[k]), "weight %s differs from the reference under the same seed" % k


def small_fixture(vae, losses, name, cfg, hw, B, T, seed):
    check_seed_parity(vae, cfg, seed)
    mats = synthetic_mats(T, 0.05, seed + 1)
    g = torch.Generator().manual_seed(seed + 2)
    lat = torch.randn(B, 4, hw, hw, generator=g)
    out = {"latents": lat.numpy(), "mats": np.packbits(mats), "T": T, "seed": seed, "strength": 5.0,
           "cfg_json": json.dumps(cfg)}
    for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        _, logits, wave, loss, grad, _ = reference_step(vae, losses, cfg, mats, lat, 5.0, dtype, seed)
        out["logits_" + tag] = logits.numpy()
        out["loss_" + tag] = loss.numpy()
        out["grad_" + tag] = grad.numpy()
    np.savez_compressed(os.path.join(GOLD, name), **out)
    print(name, "loss", out["loss_f64"], "grad norm", np.linalg.norm(out["grad_f64"]))


WFC_DATA = os.path.join(ROOT, "wavefunctiondiffusion_b200", "data", "wfc_mats")  # input data of the path: ships with the package


def wfc_mats_fixtures():
    os.makedirs(WFC_DATA, exist_ok=True)
    for path in sorted(glob.glob(REFERENCE_ROOT + "/wfd/scmap/tile_data/wfc/*.npz")):
        name = os.path.basename(path)[:-4]
        z = np.load(path)
        m = z["wfc_mats"]
        T = int(m.shape[1])
        assert T == int(z["shrink_count"])
        t_pad = (T + 127) // 128 * 128
        xs = np.argwhere(m[0]).astype(np.int16)
        ys = np.argwhere(m[1]).astype(np.int16)
        kat = (2.0 * T * T - len(xs) - len(ys)) / (2.0 * t_pad * t_pad)
        np.savez_compressed(os.path.join(WFC_DATA, name + ".npz"), T=T, T_pad=t_pad, x=xs, y=ys, uniform_kat=kat)
        print(name, T, t_pad, len(xs), len(ys), "%.8f" % kat)


def full_fixture(vae, losses):
    cfg = json.load(open(REFERENCE_ROOT + "/configs/tilenet/tilenet_sc_space_64x64.json"))
    mats = np.load(REFERENCE_ROOT + "/wfd/scmap/tile_data/wfc/ice_64x64.npz")["wfc_mats"]
    T = mats.shape[1]
    t_pad = (T + 127) // 128 * 128
    cfg["in_channels"] = cfg["out_channels"] = t_pad
    check_seed_parity(vae, cfg, 0)
    g = torch.Generator().manual_seed(1234)
    lat = torch.randn(1, 4, 64, 64, generator=g)
    out = {"latents": lat.numpy(), "seed": 0, "strength": 5.0, "cfg_json": json.dumps(cfg), "tileset": "ice_64x64"}
    pix = np.array([0, 63, 64 * 31 + 17, 64 * 63, 64 * 64 - 1, 1234, 2048, 3000])
    for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        _, logits, wave, loss, grad, dt = reference_step(vae, losses, cfg, mats, lat, 5.0, dtype, 0)
        out["loss_" + tag] = loss.numpy()
        out["grad_" + tag] = grad.numpy().astype(np.float32 if tag == "f32" else np.float64)
        out["tiles_" + tag] = wave[0].permute(1, 2, 0).numpy().argmax(axis=2).astype(np.int16)
        out["logits_pix_" + tag] = logits[0].reshape(t_pad, -1)[:, pix].t().numpy().astype(np.float32)
        lg = logits[0].reshape(t_pad, -1)
        top2 = torch.topk(lg, 2, dim=0).values
        out["top2_margin_" + tag] = (top2[0] - top2[1]).numpy().astype(np.float32)
        out["logits_mean_std_" + tag] = np.array([lg.mean().item(), lg.std().item()])
        out["seconds_" + tag] = dt
        print("ice C1", tag, "loss", loss.numpy(), "grad L2", grad.norm().item(), "absmax", grad.abs().max().item(),
              "%.1fs" % dt)
    out["pix"] = pix
    np.savez_compressed(os.path.join(GOLD, "ice_c1_step.npz"), **out)


def encode_fixture(vae):
    """AutoencoderTile.encode of the unmodified reference on a one-hot tile map and on a soft wave (tiny config)."""
    torch.manual_seed(7)
    rng = np.random.RandomState(17)
    tiles = rng.randint(0, 100, size=(2, 16, 16))
    onehot = torch.nn.functional.one_hot(torch.from_numpy(tiles), num_classes=TINY["in_channels"]).permute(0, 3, 1, 2)
    soft = torch.softmax(torch.randn(2, TINY["in_channels"], 16, 16, generator=torch.Generator().manual_seed(3)) * 2, dim=1)
    out = {"tiles": tiles, "soft": soft.numpy(), "seed": 7, "cfg_json": json.dumps(TINY)}
    for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        torch.manual_seed(7)
        net = vae.AutoencoderTile(**TINY).eval().to(dtype)
        with torch.no_grad():
            out["moments_onehot_" + tag] = net.encode(onehot.to(dtype)).latent_dist.parameters.numpy()
            out["moments_soft_" + tag] = net.encode(soft.to(dtype)).latent_dist.parameters.numpy()
    np.savez_compressed(os.path.join(GOLD, "tiny_encode.npz"), **out)
    print("tiny_encode", out["moments_onehot_f64"].shape, float(np.abs(out["moments_onehot_f64"]).mean()))


def tiles_post_fixture():
    """Post-argmax integer work on the ice tileset: the reference's own add_symmetry (wfd/scmap/data_processing.py:191-196,
    executed from where it lies) and the id chain gen_to_sc[shrink_to_gen[.]] (map_display.py:55) on the shipped tables."""
    import importlib
    import types

    for name, path in (("ref_wfd", "/wfd"), ("ref_wfd.scmap", "/wfd/scmap")):
        m = types.ModuleType(name)
        m.__path__ = [REFERENCE_ROOT + path]  # packages without running their __init__
        sys.modules[name] = m
    dp = importlib.import_module("ref_wfd.scmap.data_processing")
    consts = np.load(REFERENCE_ROOT + "/wfd/scmap/tile_data/wfc/ice_64x64.npz")
    symm_path = REFERENCE_ROOT + "/wfd/scmap/tile_data/symmetry/ice.npz"
    rng = np.random.RandomState(5)
    tiles = rng.randint(0, int(consts["shrink_count"]), size=(12, 9))
    mirrored = dp.add_symmetry(tiles, symmetry_path=symm_path)
    sc = consts["gen_to_sc"][consts["shrink_to_gen"][mirrored]]           # map_display.py:55
    np.savez_compressed(os.path.join(GOLD, "tiles_post.npz"), tiles=tiles, mirrored=mirrored, sc=sc,
                        mapping=np.load(symm_path)["mapping"], shrink_to_gen=consts["shrink_to_gen"],
                        gen_to_sc=consts["gen_to_sc"])
    print("tiles_post", mirrored.shape, sc[0, :6])


def batch8_fixture(vae, losses):
    """BASELINE config 2's shape (ice 64x64, batch 8): slot 0 carries the C1 latents, slots 1-7 are fresh draws, slot 5
    repeats slot 2 (identical maps must give identical results). loss (8,) and latent gradient, fp64 and fp32."""
    cfg = json.load(open(REFERENCE_ROOT + "/configs/tilenet/tilenet_sc_space_64x64.json"))
    mats = np.load(REFERENCE_ROOT + "/wfd/scmap/tile_data/wfc/ice_64x64.npz")["wfc_mats"]
    T = mats.shape[1]
    t_pad = (T + 127) // 128 * 128
    cfg["in_channels"] = cfg["out_channels"] = t_pad
    c1 = np.load(os.path.join(GOLD, "ice_c1_step.npz"))
    g = torch.Generator().manual_seed(4321)
    lat = torch.randn(8, 4, 64, 64, generator=g)
    lat[0] = torch.from_numpy(c1["latents"])[0]
    lat[5] = lat[2]
    out = {"latents": lat.numpy(), "seed": 0, "strength": 5.0, "cfg_json": json.dumps(cfg), "tileset": "ice_64x64"}
    for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        _, logits, wave, loss, grad, dt = reference_step(vae, losses, cfg, mats, lat, 5.0, dtype, 0)
        out["loss_" + tag] = loss.numpy().astype(np.float64)
        out["grad_" + tag] = grad.numpy().astype(np.float32)
        out["seconds_" + tag] = dt
        print("ice C2", tag, "loss", loss.numpy(), "grad L2", grad.norm().item(), "%.1fs" % dt)
    # the reference's own batch independence: slot 0 of the batch equals the batch-1 run of the same latents
    assert abs(out["loss_f64"][0] - float(np.asarray(c1["loss_f64"]).reshape(-1)[0])) < 1e-9
    np.savez_compressed(os.path.join(GOLD, "ice_c2_step.npz"), **out)


def propagate_fixture():
    """The UNMODIFIED reference solver class, imported from where it lies (it needs numpy only): constructor state, then
    its own depth-1 propagate_information (:136-174) from every allowed cell in raster order - what double_check (:176-199)
    does - repeated until the state stops changing."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("ref_wfc_prior_dist", REFERENCE_ROOT + "/wfd/wfc/wfc_prior_dist.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    Ref = mod.WFCGeneratorPriorDist
    from wavefunctiondiffusion_b200.wfc import connections_from_wfc_mats, pack_bits  # bit packing of the fixture only
    from wavefunctiondiffusion_b200.utils.tilesets import load_wfc_mats

    def closure(gen):
        rounds = 0
        while True:
            before = gen.map_allowed_tiles.copy()
            ok = True
            for y in range(gen.max_y):
                for x in range(gen.max_x):
                    if gen.is_grid_allowed((y, x)):
                        ok = gen.propagate_information((y, x), depth=1, verbose=False) and ok
            rounds += 1
            if not ok or (before == gen.map_allowed_tiles).all():
                return ok, rounds

    out = {}
    rng = np.random.RandomState(11)
    # A: rules learnt by the reference's own get_probs_from_map from a random map of 47 tiles (+ the null tile 0); every
    # window of that map satisfies them, so seeds that contain the window's tiles never contradict
    T = 48
    src = rng.randint(1, T, size=(20, 22))
    probe = Ref((4, 4), None, np.ones((T, 4, T), dtype=np.float32))
    conn_counts, _ = probe.get_probs_from_map(src)
    H, W = 10, 12
    boundary = np.zeros((H, W), dtype=bool)
    boundary[3, 4] = boundary[7, 9] = True
    gen = Ref((W, H), None, conn_counts.copy(), boundary_grids=boundary.copy())
    out["a_conn_probs"] = conn_counts
    out["a_connections"] = gen.connections.copy()
    out["a_allowed_ctor"] = gen.map_allowed_tiles.copy()
    window = src[5:5 + H, 6:6 + W]
    seeds = rng.rand(H, W, T) < 0.35
    np.put_along_axis(seeds, window[..., None], True, axis=2)
    gen.map_allowed_tiles &= seeds
    out["a_boundary"] = boundary
    out["a_allowed_in"] = gen.map_allowed_tiles.copy()
    ok, rounds = closure(gen)
    assert ok
    out["a_allowed_out"] = gen.map_allowed_tiles.copy()
    out["a_ok"] = ok
    print("propagate A: rounds", rounds, "allowed", int(out["a_allowed_in"].sum()), "->", int(out["a_allowed_out"].sum()))
    # B: the ice rules (T = 3074: 97 words per cell), 6 x 6 map, two decided cells (the first candidates, in id order, that
    # connect on all four sides and leave the map satisfiable)
    mx, my = load_wfc_mats("ice_64x64")[:2]
    conn = connections_from_wfc_mats(mx, my).astype(np.float32)
    cb = conn > 0
    cand = [int(t) for t in np.where((cb.sum(axis=2) > 3).all(axis=1))[0] if t > 0]
    for k in range(len(cand) - 1):
        gen = Ref((6, 6), None, conn.copy())
        Ti = gen.tile_count
        for cell, t in (((2, 3), cand[k]), ((4, 0), cand[k + 1])):
            gen.map_allowed_tiles[cell] = False
            gen.map_allowed_tiles[cell][t] = True
        b_in = gen.map_allowed_tiles.copy()
        ok, rounds = closure(gen)
        if ok:
            break
    assert ok
    out["b_allowed_in"] = pack_bits(b_in)
    out["b_allowed_out"] = pack_bits(gen.map_allowed_tiles)
    out["b_ok"] = ok
    print("propagate B (ice): tiles", cand[k], cand[k + 1], "rounds", rounds, "allowed", int(b_in.sum()), "->",
          int(gen.map_allowed_tiles.sum()), "of", 36 * Ti)
    # C: a contradiction - two neighbouring cells decided to tiles that the learnt rules never put side by side
    gen = Ref((5, 4), None, conn_counts.copy())
    c2 = gen.connections
    pair = next((s, t) for s in range(1, T) for t in range(1, T) if not c2[s, 2, t])
    gen.map_allowed_tiles[1, 1] = False
    gen.map_allowed_tiles[1, 1][pair[0]] = True
    gen.map_allowed_tiles[1, 2] = False
    gen.map_allowed_tiles[1, 2][pair[1]] = True
    out["c_allowed_in"] = gen.map_allowed_tiles.copy()
    ok, _ = closure(gen)
    assert not ok
    out["c_ok"] = ok
    print("propagate C: contradiction reported by the reference")
    np.savez_compressed(os.path.join(GOLD, "wfc_propagate.npz"), **out)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    if "--only-propagate" in sys.argv:
        propagate_fixture()
        sys.exit(0)
    torch.set_num_threads(os.cpu_count())
    vae, losses = load_reference()
    if "--only-encode" in sys.argv:
        encode_fixture(vae)
        sys.exit(0)
    if "--only-tiles" in sys.argv:
        tiles_post_fixture()
        sys.exit(0)
    if "--only-c2" in sys.argv:
        batch8_fixture(vae, losses)
        sys.exit(0)
    wfc_mats_fixtures()
    tiles_post_fixture()
    encode_fixture(vae)
    small_fixture(vae, losses, "tiny_step.npz", TINY, 16, 2, 100, 7)
    small_fixture(vae, losses, "tiny_down_step.npz", TINY_DOWN, 32, 2, 100, 11)
    if "--skip-full" not in sys.argv:
        full_fixture(vae, losses)
        batch8_fixture(vae, losses)
    propagate_fixture()
