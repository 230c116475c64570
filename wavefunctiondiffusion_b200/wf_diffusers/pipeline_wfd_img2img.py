"""WaveFunctionDiffusionImg2ImgPipeline — host-side mirror of the reference img2img pipeline
(wfd/wf_diffusers/pipeline_wfd_img2img.py:76-720). Its guidance block (:667-677, :688-700) is the same code as the
txt2img one and is served by the same library calls; what differs is how the initial latents are made: from an image
through the SD VAE encoder, or from an integer tile map through the tile-VAE encoder (`image_is_wave=True`,
`preprocess_wave` :68-74, `prepare_latents` :468-531). The tile-VAE encoder runs in libwfd_b200.so on a CUDA device
(SURVEY §8f-1): an integer tile map is encoded by its indices, the one-hot conv_in becoming a gather of weight columns.

Note: the reference's `set_precision("float")` calls `wfc_loss.half()` by mistake (:253); this mirror converts to float.
"""
from typing import Callable, List, Optional, Union

import numpy as np
import torch

from .pipeline_wfd import LATENT_SCALE, WaveFunctionDiffusionPipeline


def preprocess(image):
    """PIL image(s) or tensor -> float tensor (B, 3, H, W) in [-1, 1], sizes rounded down to multiples of 32."""
    if isinstance(image, torch.Tensor):
        return image
    try:
        import PIL.Image
    except ImportError:  # pragma: no cover
        PIL = None
    if PIL is not None and isinstance(image, PIL.Image.Image):
        image = [image]
    if PIL is not None and isinstance(image[0], PIL.Image.Image):
        w, h = image[0].size
        w, h = w - w % 32, h - h % 32
        arr = np.stack([np.asarray(i.resize((w, h), resample=PIL.Image.LANCZOS)) for i in image]).astype(np.float32) / 255.0
        return torch.from_numpy(2.0 * arr.transpose(0, 3, 1, 2) - 1.0)
    if isinstance(image[0], torch.Tensor):
        return torch.cat(list(image), dim=0)
    return image


def preprocess_wave(wave, num_classes):
    """Integer tile map (H, W) or (B, H, W) -> one-hot wave (B, num_classes, H, W)."""
    if not isinstance(wave, torch.Tensor):
        wave = torch.as_tensor(np.asarray(wave), dtype=torch.int64)
    if wave.dim() == 2:
        wave = wave[None, ...]
    return torch.nn.functional.one_hot(wave.long(), num_classes=num_classes).permute(0, 3, 1, 2)


class WaveFunctionDiffusionImg2ImgPipeline(WaveFunctionDiffusionPipeline):
    def get_timesteps(self, num_inference_steps, strength, device):
        init_timestep = min(int(num_inference_steps * strength), num_inference_steps)
        t_start = max(num_inference_steps - init_timestep, 0)
        return self.scheduler.timesteps[t_start:], num_inference_steps - t_start

    def prepare_latents(self, image, timestep, batch_size, num_images_per_prompt, dtype, device, generator=None,
                        image_is_wave=False):
        # an integer tile map (B, H, W) stays integer: the tile VAE encodes it by its indices (one-hot conv_in = a gather)
        by_index = image_is_wave and image.dim() == 3 and not image.is_floating_point()
        image = image.to(device=device) if by_index else image.to(device=device, dtype=dtype)
        batch_size = batch_size * num_images_per_prompt
        if isinstance(generator, list) and len(generator) != batch_size:
            raise ValueError(f"You have passed a list of generators of length {len(generator)}, but requested an"
                             f" effective batch size of {batch_size}.")
        enc = self.tile_vae if image_is_wave else self.vae
        encode = enc.encode_tiles if by_index else enc.encode
        if isinstance(generator, list):
            init = torch.cat([encode(image[i:i + 1]).latent_dist.sample(generator[i]) for i in range(batch_size)])
        else:
            init = encode(image).latent_dist.sample(generator)
        init = init.to(dtype)
        init = LATENT_SCALE * init
        if batch_size > init.shape[0]:
            if batch_size % init.shape[0] != 0:
                raise ValueError(f"Cannot duplicate `image` of batch size {init.shape[0]} to {batch_size} text prompts.")
            init = torch.cat([init] * (batch_size // init.shape[0]), dim=0)
        if isinstance(generator, list):
            noise = torch.cat([torch.randn((1,) + init.shape[1:], generator=g, device=g.device, dtype=dtype).to(device)
                               for g in generator])
        else:
            gdev = generator.device if generator is not None else device
            noise = torch.randn(init.shape, generator=generator, device=gdev, dtype=dtype).to(device)
        return self.scheduler.add_noise(init, noise, timestep)

    def check_inputs(self, prompt, strength, callback_steps):
        if not isinstance(prompt, str) and not isinstance(prompt, list):
            raise ValueError(f"`prompt` has to be of type `str` or `list` but is {type(prompt)}")
        if strength < 0 or strength > 1:
            raise ValueError(f"The value of strength should in [1.0, 1.0] but is {strength}")
        if callback_steps is None or not isinstance(callback_steps, int) or callback_steps <= 0:
            raise ValueError(f"`callback_steps` has to be a positive integer but is {callback_steps} of type"
                             f" {type(callback_steps)}.")

    @torch.no_grad()
    def __call__(
        self,
        prompt: Union[str, List[str]],
        image=None,
        strength: float = 0.8,
        num_inference_steps: Optional[int] = 50,
        guidance_scale: Optional[float] = 3.5,
        negative_prompt: Optional[Union[str, List[str]]] = None,
        num_images_per_prompt: Optional[int] = 1,
        eta: Optional[float] = 0.0,
        generator: Optional[Union[torch.Generator, List[torch.Generator]]] = None,
        output_type: Optional[str] = "pil",
        return_dict: bool = True,
        callback: Optional[Callable[[int, int, torch.FloatTensor], None]] = None,
        callback_steps: Optional[int] = 1,
        wfc_guidance_start_step: Optional[int] = 0,
        wfc_guidance_strength: Optional[float] = 1.0,
        wfc_guidance_final_steps: Optional[int] = 0,
        wfc_guidance_final_strength: Optional[float] = 1.0,
        color_balance: Optional[List[float]] = None,
        image_is_wave: bool = False,
        **kwargs,
    ):
        if image is None:
            image = kwargs.pop("init_image", None)
        self.check_inputs(prompt, strength, callback_steps)
        if color_balance is not None:
            raise NotImplementedError("color_balance is not part of the guidance path")
        batch_size = 1 if isinstance(prompt, str) else len(prompt)
        device = self._execution_device
        do_cfg = guidance_scale > 1.0
        text_embeddings = self._encode_prompt(prompt, device, num_images_per_prompt, do_cfg, negative_prompt)
        if image_is_wave:
            tiles = image if isinstance(image, torch.Tensor) else torch.as_tensor(np.asarray(image))
            if not tiles.is_floating_point() and tiles.dim() in (2, 3) and hasattr(self.tile_vae, "encode_tiles"):
                image = tiles.long() if tiles.dim() == 3 else tiles.long()[None]   # what preprocess_wave would one-hot
            else:
                image = preprocess_wave(image, num_classes=self.tile_vae.in_channels)
        else:
            image = preprocess(image)
        self.scheduler.set_timesteps(num_inference_steps, device=device)
        timesteps, num_inference_steps = self.get_timesteps(num_inference_steps, strength, device)
        latent_timestep = timesteps[:1].repeat(batch_size * num_images_per_prompt)
        self.wfc_loss.to(device)
        latents = self.prepare_latents(image, latent_timestep, batch_size, num_images_per_prompt, text_embeddings.dtype,
                                       device, generator, image_is_wave)
        extra_step_kwargs = self.prepare_extra_step_kwargs(generator, eta)
        latents = self._denoise(latents, timesteps, num_inference_steps, text_embeddings, do_cfg, guidance_scale,
                                extra_step_kwargs, callback, callback_steps, wfc_guidance_start_step,
                                wfc_guidance_strength, wfc_guidance_final_steps, wfc_guidance_final_strength)
        return self._finish(latents, device, text_embeddings.dtype, output_type, return_dict)
