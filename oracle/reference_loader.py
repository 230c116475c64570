"""TEST INFRASTRUCTURE ONLY — imports the UNMODIFIED reference modules from /root/reference for golden-vector
generation in the authoring container (the GPU box has no /root/reference, nothing at run time may call this).

`diffusers` is not installed, so the seven module names the reference files import from it are registered as inert
stand-ins (config/model mixins, BaseOutput), and the reference's own vendored UNetMidBlock2D
(wfd/wf_diffusers/wf_unet_2d_blocks.py:388-464) is aliased where wf_vae.py:24 expects the third-party one. No reference
code is copied: the files are executed from where they lie.
"""
import functools
import importlib
import importlib.util
import inspect
import os
import sys
import types
from collections import OrderedDict
from dataclasses import fields

import torch

REFERENCE_ROOT = "/root/reference"
# The one offline install of the reference (`pip install --no-deps --target baseline/_ref`): git-ignored,
# but it travels to the GPU box with the snapshot, so the UNMODIFIED reference modules can also be executed there
# (reference arm of bench.py, tests/test_patch_reference_gpu.py).
INSTALLED_ROOT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def reference_root():
    """Directory that holds the reference's `wfd/` package: the read-only tree in the authoring container, else the
    installed copy; None when neither exists."""
    for root in (REFERENCE_ROOT, INSTALLED_ROOT):
        if os.path.exists(os.path.join(root, "wfd", "wf_diffusers", "wf_vae.py")):
            return root
    return None


def load_reference(ref=None):
    """Returns (wf_vae module, losses module) of the reference."""
    if "ref_wf_diffusers.wf_vae" in sys.modules:
        return sys.modules["ref_wf_diffusers.wf_vae"], sys.modules["ref_losses"]
    ref = ref or reference_root()
    if ref is None:
        raise FileNotFoundError("the reference is neither at %s nor installed under %s" % (REFERENCE_ROOT, INSTALLED_ROOT))

    def mod(name, package=False):
        m = types.ModuleType(name)
        if package:
            m.__path__ = []
        sys.modules[name] = m
        return m

    d = mod("diffusers", True)
    cu = mod("diffusers.configuration_utils")
    ut = mod("diffusers.utils", True)
    iu = mod("diffusers.utils.import_utils")
    ms = mod("diffusers.models", True)
    ub = mod("diffusers.models.unet_2d_blocks")
    em = mod("diffusers.models.embeddings")

    class ConfigMixin:
        pass

    def register_to_config(init):
        @functools.wraps(init)
        def inner(self, *a, **kw):
            bound = inspect.signature(init).bind(self, *a, **kw)
            bound.apply_defaults()
            init(self, *a, **kw)
            self.config = types.SimpleNamespace(**{k: v for k, v in bound.arguments.items() if k != "self"})

        return inner

    class BaseOutput(OrderedDict):
        def __post_init__(self):
            for f in fields(self):
                self[f.name] = getattr(self, f.name)

    cu.ConfigMixin, cu.register_to_config, cu.FrozenDict = ConfigMixin, register_to_config, dict
    d.ModelMixin = type("ModelMixin", (torch.nn.Module,), {})
    d.configuration_utils, d.utils, d.models = cu, ut, ms
    ut.BaseOutput = BaseOutput
    ut.import_utils = iu
    iu.is_xformers_available = lambda: False
    em.ImagePositionalEmbeddings = type("ImagePositionalEmbeddings", (torch.nn.Module,), {})
    ms.unet_2d_blocks, ms.embeddings = ub, em
    pkg = mod("ref_wf_diffusers")
    pkg.__path__ = [ref + "/wfd/wf_diffusers"]  # package without running its __init__
    ub.UNetMidBlock2D = importlib.import_module("ref_wf_diffusers.wf_unet_2d_blocks").UNetMidBlock2D
    vae = importlib.import_module("ref_wf_diffusers.wf_vae")
    spec = importlib.util.spec_from_file_location("ref_losses", ref + "/wfd/wfdf/losses.py")
    losses = importlib.util.module_from_spec(spec)
    sys.modules["ref_losses"] = losses
    spec.loader.exec_module(losses)
    return vae, losses
