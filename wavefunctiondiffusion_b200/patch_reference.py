"""Keep the reference's own objects, replace only the arithmetic (INTEGRATION.md section B as a shipped shim).

    from wfd.wf_diffusers import WaveFunctionDiffusionPipeline, AutoencoderTile      # the UNMODIFIED reference
    from wavefunctiondiffusion_b200.patch_reference import patch_pipeline
    pipeline = WaveFunctionDiffusionPipeline.from_pretrained(...); pipeline.to("cuda")
    patch_pipeline(pipeline)          # tile_vae._decode and wfc_loss.forward now run in libwfd_b200.so

What is rebound, per object (nothing in the reference's modules or classes is edited):
  * `AutoencoderTile._decode` (reference wf_vae.py:592-599): `post_quant_conv` + `DecoderTile.forward` become one
    `wfd_decoder_forward` call; the data gradient comes from `wfd_decoder_backward` through a torch.autograd.Function,
    so the reference's own `torch.autograd.grad(_loss, _latents)` (pipeline_wfd.py:613) keeps working. The packed device
    weights are rebuilt from the object's `state_dict()` whenever `.to()/.half()/.float()/load_state_dict()` is called.
  * `WFCLossEinsum.forward` (reference losses.py:63-75): softmax (optional) + the four-direction adjacency contraction
    become `wfd_wfc_forward` / `wfd_wfc_backward`; the adjacency matrices are read back from the object's own `cx`/`cy`
    buffers.
`unpatch(obj)` restores the reference's methods. CPU tensors raise WfdError afterwards: there is no CPU path.
"""
import types

import torch

from .wf_diffusers.native_ops import NativeDecoder, TileDecodeFunction

_MARK = "_wfd_b200_patch"


def _is_patched(obj):
    return _MARK in vars(obj)


def patch_autoencoder_tile(model):
    """Rebinds `_decode` of ONE reference AutoencoderTile instance to the CUDA path; returns the instance."""
    if _is_patched(model):
        return model
    for name in ("latent_channels", "out_channels", "block_out_channels", "conv_groups", "up_block_types",
                 "conv_init_groups", "layers_per_block", "norm_num_groups"):
        if not hasattr(model.config, name):
            raise TypeError("patch_autoencoder_tile: %r has no config.%s - not a tile VAE" % (type(model).__name__, name))
    native = NativeDecoder(model)       # packs decoder.* / post_quant_conv.* of model.state_dict() on first use
    module = __import__(type(model).__module__, fromlist=["DecoderOutput"])
    out_cls = getattr(module, "DecoderOutput", None)
    saved = {}

    def _decode(self, z, return_dict=True):
        dec = TileDecodeFunction.apply(z, native)
        if not return_dict:
            return (dec,)
        return out_cls(sample=dec) if out_cls is not None else types.SimpleNamespace(sample=dec)

    def _apply(self, fn, *a, **kw):
        r = type(self)._apply(self, fn, *a, **kw)
        native.invalidate()
        return r

    def load_state_dict(self, *a, **kw):
        r = type(self).load_state_dict(self, *a, **kw)
        native.invalidate()
        return r

    for name, fn in (("_decode", _decode), ("_apply", _apply), ("load_state_dict", load_state_dict)):
        saved[name] = vars(model).get(name)
        setattr(model, name, types.MethodType(fn, model))
    object.__setattr__(model, _MARK, {"saved": saved, "native": native})
    return model


def patch_wfc_loss(loss):
    """Rebinds `forward` of ONE reference WFCLossEinsum instance to the CUDA path; returns the instance."""
    if _is_patched(loss):
        return loss
    if not (hasattr(loss, "cx") and hasattr(loss, "cy") and hasattr(loss, "count")):
        raise TypeError("patch_wfc_loss: %r has no cx / cy / count - not a WFCLossEinsum" % type(loss).__name__)
    from .wfdf.losses import WFCLossEinsum

    ax = (loss.cx[0, :, :, 0, 0].float() < 0.5).cpu().numpy()    # buffers hold 1 - adjacency (losses.py:56-59)
    ay = (loss.cy[0, :, :, 0, 0].float() < 0.5).cpu().numpy()
    mine = WFCLossEinsum(ax, ay)
    saved = {"forward": vars(loss).get("forward")}

    def forward(self, t, softmax=False):
        return mine(t, softmax)

    loss.forward = types.MethodType(forward, loss)
    object.__setattr__(loss, _MARK, {"saved": saved, "mirror": mine})
    return loss


def patch_pipeline(pipe):
    """Rebinds the two objects a reference pipeline (txt2img or img2img) reaches the hot path through."""
    patch_autoencoder_tile(pipe.tile_vae)
    patch_wfc_loss(pipe.wfc_loss)
    return pipe


def unpatch(obj):
    """Restores the reference's own methods on an instance patched above (or on both objects of a pipeline)."""
    if hasattr(obj, "tile_vae") and hasattr(obj, "wfc_loss") and not _is_patched(obj):
        unpatch(obj.tile_vae)
        unpatch(obj.wfc_loss)
        return obj
    rec = vars(obj).get(_MARK)
    if rec is None:
        return obj
    for name, old in rec["saved"].items():
        if old is None:
            try:
                object.__delattr__(obj, name)
            except AttributeError:
                pass
        else:
            object.__setattr__(obj, name, old)
    object.__delattr__(obj, _MARK)
    return obj
