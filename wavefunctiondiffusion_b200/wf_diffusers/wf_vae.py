"""AutoencoderTile — host-side mirror of the reference tile VAE (wfd/wf_diffusers/wf_vae.py:509-655).

Same constructor arguments, attribute names, state_dict keys/shapes and default initialisation order as the reference
(so a reference checkpoint loads unchanged and `torch.manual_seed(s)` yields the same random-init weights), but
`decode()` does not run torch modules: it goes through the C ABI of libwfd_b200.so (sm_100a kernels) and registers a
torch.autograd.Function so that `torch.autograd.grad(loss, latents)` of the guidance block (pipeline_wfd.py:613) keeps
working. `encode()` (the img2img wave input) runs in the library too when the tensor is on a CUDA device
and the encoder keeps the resolution; its eager torch modules serve the remaining cases. The decoder modules below
exist to own parameters.
"""
import json
import os
from collections import OrderedDict
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .native_ops import NativeDecoder, NativeEncoder, TileDecodeFunction


# --------------------------------------------------------------------------------------------- output containers
class _Output(OrderedDict):
    """dict + attribute access, the small part of diffusers.utils.BaseOutput user code relies on."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __getitem__(self, k):
        if isinstance(k, int):
            return list(self.values())[k]
        return super().__getitem__(k)


class DecoderOutput(_Output):
    def __init__(self, sample):
        super().__init__(sample=sample)


class AutoencoderKLOutput(_Output):
    def __init__(self, latent_dist):
        super().__init__(latent_dist=latent_dist)


class DiagonalGaussianDistribution:
    """Mean / log-variance split of the encoder moments (reference wf_vae.py:349-394)."""

    def __init__(self, parameters, deterministic=False):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.deterministic = deterministic
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)
        if deterministic:
            self.var = self.std = torch.zeros_like(self.mean)

    def sample(self, generator: Optional[torch.Generator] = None):
        noise = torch.randn(self.mean.shape, generator=generator, device="cpu" if generator is not None and generator.device.type == "cpu" else self.parameters.device)
        return self.mean + self.std * noise.to(device=self.parameters.device, dtype=self.mean.dtype)

    def mode(self):
        return self.mean

    def kl(self, other=None):
        if self.deterministic:
            return torch.zeros(1, device=self.parameters.device)
        if other is None:
            return 0.5 * torch.sum(self.mean.pow(2) + self.var - 1.0 - self.logvar, dim=[1, 2, 3])
        return 0.5 * torch.sum(
            (self.mean - other.mean).pow(2) / other.var + self.var / other.var - 1.0 - self.logvar + other.logvar,
            dim=[1, 2, 3],
        )


# --------------------------------------------------------------------------------------------- parameter containers
class _Resnet(nn.Module):
    """GN -> SiLU -> grouped 3x3 conv, twice, plus (1x1 grouped) shortcut. Reference resnet.py:369-499 with temb=None."""

    def __init__(self, cin, cout, conv_groups, norm_groups):
        super().__init__()
        self.norm1 = nn.GroupNorm(norm_groups, cin, eps=1e-6, affine=True)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1, groups=conv_groups)
        self.norm2 = nn.GroupNorm(norm_groups, cout, eps=1e-6, affine=True)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1, groups=conv_groups)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1, groups=conv_groups) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        return (x if self.conv_shortcut is None else self.conv_shortcut(x)) + h


class _Attention(nn.Module):
    """Single-head spatial self-attention (reference attention.py:247-380)."""

    def __init__(self, channels, norm_groups):
        super().__init__()
        self.channels = channels
        self.group_norm = nn.GroupNorm(norm_groups, channels, eps=1e-6, affine=True)
        self.query = nn.Linear(channels, channels)
        self.key = nn.Linear(channels, channels)
        self.value = nn.Linear(channels, channels)
        self.proj_attn = nn.Linear(channels, channels)

    def forward(self, x):
        b, c, hh, ww = x.shape
        t = self.group_norm(x).view(b, c, hh * ww).transpose(1, 2)
        q, k, v = self.query(t), self.key(t), self.value(t)
        p = torch.softmax((q @ k.transpose(1, 2)).float() * (c ** -0.5), dim=-1).to(q.dtype)
        o = self.proj_attn(p @ v).transpose(1, 2).reshape(b, c, hh, ww)
        return o + x


class _MidBlock(nn.Module):
    def __init__(self, channels, norm_groups):
        super().__init__()
        r0 = _Resnet(channels, channels, 1, norm_groups)
        a0 = _Attention(channels, norm_groups)
        r1 = _Resnet(channels, channels, 1, norm_groups)
        self.attentions = nn.ModuleList([a0])
        self.resnets = nn.ModuleList([r0, r1])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class _Resample(nn.Module):
    """`.conv` holder of Downsample2D(name='op') / Upsample2D (reference resnet.py:77-192)."""

    def __init__(self, channels, conv_groups, down):
        super().__init__()
        self.down = down
        if down:
            self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=0, groups=conv_groups)
        else:
            self.conv = nn.Conv2d(channels, channels, 3, padding=1, groups=conv_groups)

    def forward(self, x):
        if self.down:
            return self.conv(F.pad(x, (0, 1, 0, 1)))
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class _Block(nn.Module):
    def __init__(self, cin, cout, n_layers, conv_groups, norm_groups, resample=None):
        super().__init__()
        self.resnets = nn.ModuleList(
            [_Resnet(cin if i == 0 else cout, cout, conv_groups, norm_groups) for i in range(n_layers)]
        )
        if resample == "down":
            self.downsamplers = nn.ModuleList([_Resample(cout, conv_groups, True)])
        elif resample == "up":
            self.upsamplers = nn.ModuleList([_Resample(cout, conv_groups, False)])

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        for name in ("downsamplers", "upsamplers"):
            for m in getattr(self, name, []):
                x = m(x)
        return x


_ENC_RESAMPLE = {"EvenEncoderBlock2D": None, "UpEncoderBlock2D": "up", "DownEncoderBlock2D": "down"}
_DEC_RESAMPLE = {"EvenDecoderBlock2D": None, "DownDecoderBlock2D": "down"}


class EncoderTile(nn.Module):
    """Reference wf_vae.py:68-148. Eager torch; not on the guidance path."""

    def __init__(self, in_channels, out_channels, down_block_types, block_out_channels, layers_per_block,
                 norm_num_groups, conv_groups, conv_init_groups, double_z=True):
        super().__init__()
        self.conv_in = nn.Conv2d(in_channels, block_out_channels[0], 3, padding=1, groups=conv_init_groups)
        blocks = []
        ch = block_out_channels[0]
        for i, kind in enumerate(down_block_types):
            if kind not in _ENC_RESAMPLE:
                raise ValueError(f"{kind} does not exist.")
            resample = _ENC_RESAMPLE[kind]
            if kind == "DownEncoderBlock2D" and i == len(block_out_channels) - 1:
                resample = None
            blocks.append(_Block(ch, block_out_channels[i], layers_per_block, conv_groups[i], norm_num_groups, resample))
            ch = block_out_channels[i]
        self.down_blocks = nn.ModuleList(blocks)
        self.mid_block = _MidBlock(block_out_channels[-1], norm_num_groups)
        self.conv_norm_out = nn.GroupNorm(norm_num_groups, block_out_channels[-1], eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(block_out_channels[-1], 2 * out_channels if double_z else out_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(self.conv_act(self.conv_norm_out(x)))


class DecoderTile(nn.Module):
    """Parameter container for reference wf_vae.py:151-232; the arithmetic lives in libwfd_b200.so."""

    def __init__(self, in_channels, out_channels, up_block_types, block_out_channels, layers_per_block, norm_num_groups,
                 conv_groups, conv_init_groups):
        super().__init__()
        rev = list(reversed(block_out_channels))
        rev_groups = list(reversed(conv_groups))
        self.conv_in = nn.Conv2d(in_channels, rev[0], 3, padding=1)
        self.up_blocks = nn.ModuleList([])  # registered before mid_block, as in the reference's state_dict order
        self.mid_block = _MidBlock(rev[0], norm_num_groups)
        ch = rev[0]
        for i, kind in enumerate(up_block_types):
            if kind not in _DEC_RESAMPLE:
                raise ValueError(f"{kind} does not exist.")
            self.up_blocks.append(
                _Block(ch, rev[i], layers_per_block + 1, rev_groups[i], norm_num_groups, _DEC_RESAMPLE[kind])
            )
            ch = rev[i]
        self.conv_norm_out = nn.GroupNorm(norm_num_groups, block_out_channels[0], eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(block_out_channels[0], out_channels, 3, padding=1, groups=conv_init_groups)

    def forward(self, z):
        raise RuntimeError(
            "DecoderTile has no eager forward in wavefunctiondiffusion_b200; use AutoencoderTile.decode (CUDA path)"
        )


# --------------------------------------------------------------------------------------------- the model
class AutoencoderTile(nn.Module):
    """Drop-in for the reference AutoencoderTile (wf_vae.py:509-655)."""

    config_name = "config.json"
    weights_name = "diffusion_pytorch_model.bin"

    def __init__(
        self,
        in_channels: int = 3,
        out_channels: int = 3,
        down_block_types: Tuple[str] = ("DownEncoderBlock2D",),
        up_block_types: Tuple[str] = ("UpDecoderBlock2D",),
        block_out_channels: Tuple[int] = (64,),
        layers_per_block: int = 1,
        act_fn: str = "silu",
        latent_channels: int = 4,
        norm_num_groups: int = 32,
        sample_size: int = 32,
        conv_groups: Tuple[int] = (1,),
        conv_init_groups: int = 1,
    ):
        super().__init__()
        if act_fn not in ("silu", "swish"):
            raise NotImplementedError(f"act_fn={act_fn!r}: the tile VAE kernels implement SiLU only")
        self.config = SimpleNamespace(
            in_channels=in_channels, out_channels=out_channels, down_block_types=tuple(down_block_types),
            up_block_types=tuple(up_block_types), block_out_channels=tuple(block_out_channels),
            layers_per_block=layers_per_block, act_fn=act_fn, latent_channels=latent_channels,
            norm_num_groups=norm_num_groups, sample_size=sample_size, conv_groups=tuple(conv_groups),
            conv_init_groups=conv_init_groups,
        )
        self.in_channels = in_channels
        self.encoder = EncoderTile(in_channels, latent_channels, down_block_types, block_out_channels, layers_per_block,
                                   norm_num_groups, conv_groups, conv_init_groups, double_z=True)
        self.decoder = DecoderTile(latent_channels, out_channels, up_block_types, block_out_channels, layers_per_block,
                                   norm_num_groups, conv_groups, conv_init_groups)
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1)
        self.post_quant_conv = nn.Conv2d(latent_channels, latent_channels, 1)
        self.use_slicing = False
        self._native = NativeDecoder(self)
        self._native_enc = NativeEncoder(self)

    # ---- diffusers-like persistence (config.json + diffusion_pytorch_model.bin)
    @classmethod
    def from_config(cls, config, **overrides):
        cfg = dict(config if isinstance(config, dict) else vars(config))
        cfg = {k: v for k, v in cfg.items() if not k.startswith("_")}
        cfg.update(overrides)
        return cls(**cfg)

    @classmethod
    def from_pretrained(cls, path, subfolder=None, **overrides):
        root = os.path.join(path, subfolder) if subfolder else path
        with open(os.path.join(root, cls.config_name)) as f:
            model = cls.from_config(json.load(f), **overrides)
        wpath = os.path.join(root, cls.weights_name)
        if os.path.exists(wpath):
            state = torch.load(wpath, map_location="cpu")
        else:
            from safetensors.torch import load_file  # optional dependency

            state = load_file(os.path.join(root, "diffusion_pytorch_model.safetensors"))
        model.load_state_dict(state)
        return model.eval()

    def save_pretrained(self, path):
        os.makedirs(path, exist_ok=True)
        with open(os.path.join(path, self.config_name), "w") as f:
            cfg = {k: (list(v) if isinstance(v, tuple) else v) for k, v in vars(self.config).items()}
            cfg["_class_name"] = "AutoencoderTile"
            json.dump(cfg, f, indent=2)
        torch.save(self.state_dict(), os.path.join(path, self.weights_name))

    # ---- weight-change tracking for the packed native copy
    def load_state_dict(self, *a, **kw):
        r = super().load_state_dict(*a, **kw)
        self._native.invalidate()
        self._native_enc.invalidate()
        return r

    def _apply(self, fn, *a, **kw):
        r = super()._apply(fn, *a, **kw)
        if hasattr(self, "_native"):
            self._native.invalidate()
        if hasattr(self, "_native_enc"):
            self._native_enc.invalidate()
        return r

    # ---- copies and pickles carry parameters only; the native caches (ctypes handles, device workspaces, a weak
    # reference to the owner) are rebuilt for the new object on its first decode
    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop("_native", None)
        state.pop("_native_enc", None)
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._native = NativeDecoder(self)
        self._native_enc = NativeEncoder(self)

    def __deepcopy__(self, memo):
        import copy

        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        new.__setstate__({k: copy.deepcopy(v, memo) for k, v in self.__getstate__().items()})
        return new

    def refresh_native(self):
        """Call after mutating weights in place; the packed device copies are rebuilt on the next decode / encode."""
        self._native.invalidate()
        self._native_enc.invalidate()

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    # ---- reference API
    def _encode_moments(self, x):
        """quant_conv(encoder(x)) (reference :583-584). On a CUDA device and for configurations whose encoder blocks keep
        the resolution this runs in libwfd_b200.so (forward only: the moments carry no autograd graph); the eager torch
        modules remain for the other cases (CPU tensors, the 32x32 family's upsampling encoder, training)."""
        native_ok = (x.is_cuda and NativeEncoder.supported(self.config) and not (torch.is_grad_enabled() and x.requires_grad)
                     and x.shape[2] * x.shape[3] % 128 == 0)
        if native_ok:
            plan = self._native_enc.plan(x.shape[0], x.shape[2], x.shape[3], x.device)
            return plan.encode_wave(x).to(x.dtype if x.dtype.is_floating_point else torch.float32)
        return self.quant_conv(self.encoder(x))

    def encode_tiles(self, tiles, return_dict: bool = True):
        """encode(preprocess_wave(tiles)) for an integer tile map (B, H, W) or (H, W) given by its indices (the img2img
        wave input, reference pipeline_wfd_img2img.py:68-74, 478-486): on the device the one-hot conv_in is a gather of
        weight columns, so the (B, T_pad, H, W) one-hot tensor is never materialised."""
        t = tiles if tiles.dim() == 3 else tiles.unsqueeze(0)
        if t.is_cuda and NativeEncoder.supported(self.config) and t.shape[1] * t.shape[2] % 128 == 0:
            moments = self._native_enc.plan(t.shape[0], t.shape[1], t.shape[2], t.device).encode_tiles(t).to(self.dtype)
        else:
            onehot = F.one_hot(t.long(), num_classes=self.in_channels).permute(0, 3, 1, 2).to(self.dtype)
            moments = self._encode_moments(onehot)
        posterior = DiagonalGaussianDistribution(moments)
        if not return_dict:
            return (posterior,)
        return AutoencoderKLOutput(latent_dist=posterior)

    def encode(self, x, return_dict: bool = True):
        moments = self._encode_moments(x)
        posterior = DiagonalGaussianDistribution(moments)
        if not return_dict:
            return (posterior,)
        return AutoencoderKLOutput(latent_dist=posterior)

    def _decode(self, z, return_dict: bool = True):
        dec = TileDecodeFunction.apply(z, self._native)
        if not return_dict:
            return (dec,)
        return DecoderOutput(sample=dec)

    def enable_slicing(self):
        self.use_slicing = True

    def disable_slicing(self):
        self.use_slicing = False

    def decode(self, z, return_dict: bool = True):
        if self.use_slicing and z.shape[0] > 1:
            decoded = torch.cat([self._decode(z_slice).sample for z_slice in z.split(1)])
        else:
            decoded = self._decode(z).sample
        if not return_dict:
            return (decoded,)
        return DecoderOutput(sample=decoded)

    def forward(self, sample, sample_posterior: bool = False, return_dict: bool = True, generator=None):
        posterior = self.encode(sample).latent_dist
        z = posterior.sample(generator=generator) if sample_posterior else posterior.mode()
        dec = self.decode(z).sample
        if not return_dict:
            return (dec,)
        return DecoderOutput(sample=dec)
