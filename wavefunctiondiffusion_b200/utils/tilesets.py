"""Tileset adjacency data: the `wfc_mats` arrays of the reference's wfd/scmap/tile_data/wfc/*.npz
(written by wfd/wfdf/datasets.py:124-146, read by pipeline_wfd.py:180-182).

`load_wfc_mats` accepts either a reference npz (key `wfc_mats`, bool (2, T, T)) or the compact coordinate form shipped
with this package under data/wfc_mats/ (keys T, x, y: the same matrices as the reference's nine npz files, bit for bit,
written by oracle/make_golden.py) so that the tileset names resolve where the reference tree is absent.
"""
import os

import numpy as np

WFC_DATA_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "wfc_mats")


def round_up_channels(t, multiple=128):
    """in/out channel rule of the tile VAE (reference train_tile_vae.py:23-24, 141-144)."""
    return (int(t) + multiple - 1) // multiple * multiple


def load_wfc_mats(path_or_name):
    """-> (mat_x, mat_y) bool (T, T). `path_or_name`: an npz path, or a tileset name like 'ice_64x64'."""
    path = path_or_name
    if not os.path.exists(path):
        path = os.path.join(WFC_DATA_DIR, path_or_name + ".npz")
    if not os.path.exists(path):
        raise FileNotFoundError("no wfc data for %r" % (path_or_name,))
    z = np.load(path)
    if "wfc_mats" in z.files:
        m = z["wfc_mats"]
        return np.asarray(m[0], dtype=bool), np.asarray(m[1], dtype=bool)
    T = int(z["T"])
    mx = np.zeros((T, T), dtype=bool)
    my = np.zeros((T, T), dtype=bool)
    xs, ys = z["x"].astype(np.int64), z["y"].astype(np.int64)
    mx[xs[:, 0], xs[:, 1]] = True
    my[ys[:, 0], ys[:, 1]] = True
    return mx, my


TILENET_64 = dict(
    down_block_types=("EvenEncoderBlock2D",) * 4, up_block_types=("EvenDecoderBlock2D",) * 4,
    block_out_channels=(3072, 1536, 768, 384), conv_groups=(64, 16, 4, 1), conv_init_groups=128, layers_per_block=2,
    act_fn="silu", latent_channels=4, norm_num_groups=8, sample_size=64,
)
TILENET_32 = dict(
    TILENET_64,
    down_block_types=("EvenEncoderBlock2D", "UpEncoderBlock2D", "EvenEncoderBlock2D", "EvenEncoderBlock2D"),
    up_block_types=("EvenDecoderBlock2D", "EvenDecoderBlock2D", "DownDecoderBlock2D", "EvenDecoderBlock2D"),
    sample_size=32,
)


def tilenet_config(name, t_pad):
    """The two shipped architectures (reference configs/tilenet/tilenet_sc_space_{64x64,32x32}.json) with the
    in/out channel count filled in."""
    base = {"64x64": TILENET_64, "32x32": TILENET_32}[name]
    return dict(base, in_channels=t_pad, out_channels=t_pad)
