"""Stand-ins for the third-party stages around the guidance path (UNet, text encoder, image VAE, scheduler), used by
the tests and by bench.py where `diffusers`/`transformers` and pretrained weights are unavailable. They follow the
call conventions the pipeline relies on; none of them is on the measured path except `AffineDDIMScheduler.step`, the
affine map of the latents that the reference's guidance block differentiates through (pipeline_wfd.py:611).
"""
import math
from types import SimpleNamespace

import numpy as np
import torch


class AffineDDIMScheduler:
    """DDIM (eta = 0) on SD's scaled-linear betas written as prev = a_i * x + b_i * eps."""

    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012):
        betas = np.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=np.float64) ** 2
        self.alphas_cumprod = np.cumprod(1.0 - betas)
        self.num_train_timesteps = num_train_timesteps
        self.timesteps = None

    def set_timesteps(self, num_inference_steps, device=None):
        self.step_size = self.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * self.step_size)[::-1].copy()
        self.timesteps = torch.from_numpy(ts).to(device)

    def scale_model_input(self, sample, t):
        return sample

    def coefficients(self, t):
        t = int(t)
        prev_t = t - self.step_size
        at = self.alphas_cumprod[t]
        ap = self.alphas_cumprod[prev_t] if prev_t >= 0 else 1.0
        return math.sqrt(ap / at), math.sqrt(1 - ap) - math.sqrt(ap * (1 - at) / at)

    def add_noise(self, original, noise, timesteps):
        acp = torch.as_tensor(self.alphas_cumprod, dtype=torch.float32, device=original.device)[timesteps.long()]
        a = acp.sqrt().view(-1, 1, 1, 1).to(original.dtype)
        b = (1 - acp).sqrt().view(-1, 1, 1, 1).to(original.dtype)
        return a * original + b * noise

    def step(self, model_output, timestep, sample, **kwargs):
        a, b = self.coefficients(timestep)
        return SimpleNamespace(prev_sample=a * sample + b * model_output)


class RandomNoiseUNet(torch.nn.Module):
    """Returns seeded Gaussian noise of the right shape: the UNet is third-party and outside the measured path."""

    def __init__(self, in_channels=4, sample_size=64):
        super().__init__()
        self.in_channels = in_channels
        self.config = SimpleNamespace(sample_size=sample_size, in_channels=in_channels)
        self.dummy = torch.nn.Parameter(torch.zeros(1))
        self._gen = None

    def forward(self, sample, t, encoder_hidden_states=None):
        if self._gen is None or self._gen.device != sample.device:
            self._gen = torch.Generator(device=sample.device).manual_seed(4321)
        noise = torch.randn(sample.shape, generator=self._gen, device=sample.device, dtype=sample.dtype)
        return SimpleNamespace(sample=noise)


class ConstantTextEncoder(torch.nn.Module):
    def __init__(self, dim=8):
        super().__init__()
        self.config = SimpleNamespace(use_attention_mask=False)
        self.emb = torch.nn.Embedding(16, dim)

    def forward(self, input_ids, attention_mask=None):
        return (self.emb(input_ids.clamp(max=15)),)


class WhitespaceTokenizer:
    model_max_length = 8

    def __call__(self, text, padding=None, max_length=None, truncation=None, return_tensors=None):
        if isinstance(text, str):
            text = [text]
        n = max_length or self.model_max_length
        ids = torch.zeros(len(text), n, dtype=torch.long)
        for i, s in enumerate(text):
            for j, w in enumerate(s.split()[:n]):
                ids[i, j] = 1 + (sum(map(ord, w)) % 15)
        return SimpleNamespace(input_ids=ids, attention_mask=(ids > 0).long())


class TinyImageVAE(torch.nn.Module):
    """Image VAE stand-in: nearest upsampling of three latent channels."""

    def __init__(self):
        super().__init__()
        self.config = SimpleNamespace(block_out_channels=(1, 1, 1, 1))
        self.use_slicing = False
        self.dummy = torch.nn.Parameter(torch.zeros(1))

    def enable_slicing(self):
        self.use_slicing = True

    def disable_slicing(self):
        self.use_slicing = False

    def decode(self, z):
        img = torch.nn.functional.interpolate(z[:, :3].float(), scale_factor=8.0, mode="nearest")
        return SimpleNamespace(sample=img.to(z.dtype))
