"""CPU: the reference's canonical entry points - `from wfd.wf_diffusers import ...` names, and
`WaveFunctionDiffusionPipeline.from_pretrained(path, tile_vae=..., wfc_data_path=...)` (reference README.md:60-64,
wfd/webui/ui.py:188-192) - on a pipeline folder whose third-party components come from a stand-in library."""
import json
import os
import sys
import types

import numpy as np
import pytest
import torch


class _Part(torch.nn.Module):
    loaded_from = None

    def __init__(self):
        super().__init__()
        self.config = types.SimpleNamespace(block_out_channels=(1, 2, 3, 4), sample_size=8)

    @classmethod
    def from_pretrained(cls, path, **kw):
        obj = cls()
        obj.loaded_from = path
        obj.kw = kw
        return obj


def _fake_library(monkeypatch):
    lib = types.ModuleType("wfd_standin_lib")
    for name in ("AutoencoderKL", "CLIPTextModel", "CLIPTokenizer", "UNet2DConditionModel", "PNDMScheduler"):
        setattr(lib, name, type(name, (_Part,), {}))
    monkeypatch.setitem(sys.modules, "wfd_standin_lib", lib)
    return lib


def _write_pipeline(tmp_path, with_tile_vae):
    from wavefunctiondiffusion_b200 import AutoencoderTile
    from wavefunctiondiffusion_b200.utils.tilesets import tilenet_config

    index = {"_class_name": "WaveFunctionDiffusionPipeline", "vae": ["wfd_standin_lib", "AutoencoderKL"],
             "text_encoder": ["wfd_standin_lib", "CLIPTextModel"], "tokenizer": ["wfd_standin_lib", "CLIPTokenizer"],
             "unet": ["wfd_standin_lib", "UNet2DConditionModel"], "scheduler": ["wfd_standin_lib", "PNDMScheduler"],
             "safety_checker": [None, None], "feature_extractor": [None, None]}
    cfg = tilenet_config("64x64", 128)
    cfg.update(block_out_channels=(128, 64, 64, 64), conv_groups=(2, 1, 1, 1), conv_init_groups=16)
    torch.manual_seed(0)
    net = AutoencoderTile(**cfg)
    if with_tile_vae:
        index["tile_vae"] = ["wfd", "AutoencoderTile"]
        net.save_pretrained(str(tmp_path / "tile_vae"))
    (tmp_path / "model_index.json").write_text(json.dumps(index))
    rng = np.random.RandomState(0)
    npz = str(tmp_path / "wfc.npz")
    np.savez(npz, wfc_mats=rng.rand(2, 100, 100) < 0.1)
    return net, npz


def test_package_exports_the_reference_names():
    import wavefunctiondiffusion_b200 as pkg
    from wavefunctiondiffusion_b200.wf_diffusers import (AutoencoderTile, WaveFunctionDiffusionImg2ImgPipeline,  # noqa: F401
                                                         WaveFunctionDiffusionPipeline, preprocess_img, preprocess_wave)
    from wavefunctiondiffusion_b200.wfdf import WFCLossEinsum  # noqa: F401

    for name in ("WaveFunctionDiffusionPipeline", "WaveFunctionDiffusionImg2ImgPipeline", "AutoencoderTile", "preprocess_wave",
                 "preprocess_img", "WFCLossEinsum"):
        assert hasattr(pkg, name)
    assert preprocess_wave(np.zeros((4, 4), dtype=np.int64), 8).shape == (1, 8, 4, 4)


def test_from_pretrained_as_the_readme_calls_it(tmp_path, monkeypatch):
    from wavefunctiondiffusion_b200 import WaveFunctionDiffusionImg2ImgPipeline, WaveFunctionDiffusionPipeline

    lib = _fake_library(monkeypatch)
    tilenet, npz = _write_pipeline(tmp_path, with_tile_vae=False)
    pipe = WaveFunctionDiffusionPipeline.from_pretrained(str(tmp_path), tile_vae=tilenet, wfc_data_path=npz)
    assert pipe.tile_vae is tilenet and isinstance(pipe.unet, lib.UNet2DConditionModel)
    assert pipe.unet.loaded_from == str(tmp_path / "unet") and pipe.safety_checker is None
    assert pipe.wfc_loss.count == 100 and pipe.vae_scale_factor == 8
    # a component handed in wins over the index; the img2img twin loads the same way
    own = lib.PNDMScheduler()
    pipe2 = WaveFunctionDiffusionImg2ImgPipeline.from_pretrained(str(tmp_path), tile_vae=tilenet, wfc_data_path=npz, scheduler=own)
    assert pipe2.scheduler is own
    with pytest.raises(ValueError):
        WaveFunctionDiffusionPipeline.from_pretrained(str(tmp_path), tile_vae=tilenet)            # no wfc_data_path
    with pytest.raises(ValueError):
        WaveFunctionDiffusionPipeline.from_pretrained(str(tmp_path), wfc_data_path=npz)            # no tile VAE anywhere
    with pytest.raises(TypeError):
        WaveFunctionDiffusionPipeline.from_pretrained(str(tmp_path), tile_vae=tilenet, wfc_data_path=npz, bogus=1)


def test_from_pretrained_loads_a_tile_vae_named_in_the_index(tmp_path, monkeypatch):
    from wavefunctiondiffusion_b200 import AutoencoderTile, WaveFunctionDiffusionPipeline

    _fake_library(monkeypatch)
    tilenet, npz = _write_pipeline(tmp_path, with_tile_vae=True)
    pipe = WaveFunctionDiffusionPipeline.from_pretrained(str(tmp_path), wfc_data_path=npz)
    assert isinstance(pipe.tile_vae, AutoencoderTile)
    a, b = tilenet.state_dict(), pipe.tile_vae.state_dict()
    assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)
    missing = json.loads((tmp_path / "model_index.json").read_text())
    missing["unet"] = ["a_library_that_is_not_installed", "UNet2DConditionModel"]
    (tmp_path / "model_index.json").write_text(json.dumps(missing))
    with pytest.raises(ImportError):
        WaveFunctionDiffusionPipeline.from_pretrained(str(tmp_path), wfc_data_path=npz)
