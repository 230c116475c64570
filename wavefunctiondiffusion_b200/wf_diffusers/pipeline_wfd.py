"""WaveFunctionDiffusionPipeline — host-side mirror of the reference txt2img pipeline
(wfd/wf_diffusers/pipeline_wfd.py:47-659) with the WFC-guidance block running on libwfd_b200.so.

What is ours: `decode_latents_tile` (:402-408), the guidance block (:605-615 and its final-stage twin :626-639),
`set_precision` (:204-220), the slicing toggles (:186-202) and the wave output (:647). Everything else the reference
pipeline touches — CLIP text encoder, SD UNet, SD image VAE, scheduler, safety checker — is third-party and is used
exactly as handed in (duck typed; `diffusers`/`transformers` are not imported here, so the class also works with the
stand-in components the tests and the benchmark use).

`__call__` keeps the reference's keyword set, including `wfc_guidance_start_step`, `wfc_guidance_strength`,
`wfc_guidance_final_steps`, `wfc_guidance_final_strength`, and returns `WaveFunctionDiffusionPipelineOutput(images, waves,
nsfw_content_detected)` with `waves` an fp32 ndarray (B, H, W, T_pad).

Difference kept on purpose: the reference calls `torch.autograd.grad(_loss, _latents)` on a (B,) loss, which raises for
B > 1; here the loss is summed over the batch first (maps are independent, so B = 1 is unchanged).
"""
import inspect
from dataclasses import dataclass
from typing import Callable, List, Optional, Union

import numpy as np
import torch

from ..wfdf.losses import WFCLossEinsum
from .native_ops import GuidanceLossFunction, cuda_graphs_enabled, guidance_loop, max_micro_batch

LATENT_SCALE = 0.18215


@dataclass
class WaveFunctionDiffusionPipelineOutput:
    """images: PIL images or ndarray (B, H, W, 3); waves: ndarray (B, H_tiles, W_tiles, T_pad) fp32;
    nsfw_content_detected: list of bool or None (reference pipeline_output.py:29-48)."""

    images: Union[list, np.ndarray]
    waves: np.ndarray
    nsfw_content_detected: Optional[List[bool]]


class WaveFunctionDiffusionPipeline:
    def __init__(self, vae, text_encoder, tokenizer, tile_vae, unet, scheduler, safety_checker=None,
                 feature_extractor=None, wfc_data_path: str = None, requires_safety_checker: bool = True):
        self.vae, self.text_encoder, self.tokenizer = vae, text_encoder, tokenizer
        self.tile_vae, self.unet, self.scheduler = tile_vae, unet, scheduler
        self.safety_checker, self.feature_extractor = safety_checker, feature_extractor
        block_out = getattr(getattr(vae, "config", None), "block_out_channels", (0, 0, 0, 0))
        self.vae_scale_factor = 2 ** (len(block_out) - 1)
        # tileset data (reference :180-182)
        self.wfc_consts = np.load(wfc_data_path)
        wfc_mat_x, wfc_mat_y = self.wfc_consts["wfc_mats"]
        self.wfc_loss = WFCLossEinsum(wfc_mat_x, wfc_mat_y)
        self.config = {"wfc_data_path": wfc_data_path, "requires_safety_checker": requires_safety_checker}

    # ------------------------------------------------------------------ loading (reference README.md:60-64, ui.py:188-192)
    COMPONENTS = ("vae", "text_encoder", "tokenizer", "tile_vae", "unet", "scheduler", "safety_checker", "feature_extractor")

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, tile_vae=None, wfc_data_path=None, **kwargs):
        """`WaveFunctionDiffusionPipeline.from_pretrained(path, tile_vae=tilenet, wfc_data_path=npz)` as the reference's
        README and web UI call it. The reference inherits this from `diffusers.DiffusionPipeline`; here the pipeline
        folder's `model_index.json` is read directly: every component that is not handed in as a keyword is loaded with
        `<library>.<class>.from_pretrained(path/<name>)` from the library the index names (`diffusers`, `transformers` -
        third-party stages, used as they are), the tile VAE by this package's AutoencoderTile. A Hugging Face hub id is
        resolved with `huggingface_hub.snapshot_download` when that package is importable."""
        import importlib
        import json
        import os

        root = pretrained_model_name_or_path
        if not os.path.isdir(root):
            try:
                from huggingface_hub import snapshot_download
            except ImportError as e:
                raise FileNotFoundError(f"{root!r} is not a local pipeline folder and huggingface_hub is not installed") from e
            root = snapshot_download(root)
        with open(os.path.join(root, "model_index.json")) as f:
            index = json.load(f)
        if wfc_data_path is None:
            wfc_data_path = index.get("wfc_data_path")
        if wfc_data_path is None:
            raise ValueError("wfc_data_path is required (the tileset's wfc/*.npz, reference pipeline_wfd.py:180-182)")
        torch_dtype = kwargs.pop("torch_dtype", None)
        parts = {}
        for name in cls.COMPONENTS:
            if name == "tile_vae" and tile_vae is not None:
                parts[name] = tile_vae
                continue
            if name in kwargs:
                parts[name] = kwargs.pop(name)
                continue
            spec = index.get(name)
            if not spec or spec[0] is None or spec[1] is None:
                if name in ("safety_checker", "feature_extractor"):
                    parts[name] = None
                    continue
                raise ValueError(f"model_index.json of {root} names no {name!r} and none was passed")
            lib_name, class_name = spec
            sub = os.path.join(root, name)
            if class_name == "AutoencoderTile":
                from .wf_vae import AutoencoderTile

                parts[name] = AutoencoderTile.from_pretrained(sub)
                continue
            try:
                module = importlib.import_module(lib_name)
            except ImportError as e:
                raise ImportError(f"component {name!r} of {root} needs the {lib_name!r} package ({class_name}); install it "
                                  f"or pass {name}=... to from_pretrained") from e
            klass = getattr(module, class_name)
            parts[name] = klass.from_pretrained(sub, torch_dtype=torch_dtype) if torch_dtype is not None else klass.from_pretrained(sub)
        requires = kwargs.pop("requires_safety_checker", index.get("requires_safety_checker", True))
        if kwargs:
            raise TypeError(f"from_pretrained got unexpected keyword arguments {sorted(kwargs)}")
        return cls(parts["vae"], parts["text_encoder"], parts["tokenizer"], parts["tile_vae"], parts["unet"],
                   parts["scheduler"], parts["safety_checker"], parts["feature_extractor"], wfc_data_path=wfc_data_path,
                   requires_safety_checker=requires)

    # ------------------------------------------------------------------ plumbing shared with diffusers pipelines
    @property
    def device(self):
        for m in (self.unet, self.tile_vae, self.vae, self.text_encoder):
            if isinstance(m, torch.nn.Module):
                try:
                    return next(m.parameters()).device
                except StopIteration:
                    continue
        return torch.device("cpu")

    _execution_device = device

    def to(self, device):
        for name in ("vae", "text_encoder", "tile_vae", "unet", "safety_checker"):
            m = getattr(self, name)
            if isinstance(m, torch.nn.Module):
                m.to(device)
        self.wfc_loss.to(device)
        return self

    def progress_bar(self, iterable=None, total=None):
        from tqdm.auto import tqdm

        cfg = getattr(self, "_progress_bar_config", {})
        return tqdm(iterable, **cfg) if iterable is not None else tqdm(total=total, **cfg)

    def set_progress_bar_config(self, **kwargs):
        self._progress_bar_config = kwargs

    @staticmethod
    def numpy_to_pil(images):
        from PIL import Image

        if images.ndim == 3:
            images = images[None, ...]
        images = (images * 255).round().astype("uint8")
        return [Image.fromarray(im) for im in images]

    # ------------------------------------------------------------------ toggles (reference :186-220)
    def enable_vae_slicing(self):
        self.vae.enable_slicing()
        self.tile_vae.enable_slicing()

    def disable_vae_slicing(self):
        self.vae.disable_slicing()
        self.tile_vae.disable_slicing()

    def set_precision(self, prec):
        if self.device == torch.device("cpu"):
            return
        if prec == "half":
            for m in (self.vae, self.tile_vae, self.unet, self.text_encoder, self.wfc_loss):
                m.half()
        elif prec == "float":
            for m in (self.vae, self.tile_vae, self.unet, self.text_encoder, self.wfc_loss):
                m.float()
        else:
            raise NotImplementedError

    # ------------------------------------------------------------------ the hot path
    def decode_latents_tile(self, latents, numpy=False):
        """latents -> wave = softmax over all T_pad channels of the tile-VAE decode (reference :402-408)."""
        latents = 1 / LATENT_SCALE * latents
        image = self.tile_vae.decode(latents).sample
        image = image.softmax(axis=1)
        if numpy:
            return image.cpu().permute(0, 2, 3, 1).float().numpy()
        return image

    def enable_row_bands(self, comm, graph=False):
        """Compute the guidance of ONE large map in row bands over the ranks of `comm` (rowband.BandComm): every rank
        runs this pipeline on the same latents (the UNet and the schedulers stay replicated, as in the reference) and
        the tile-VAE decode / WFC loss / backward of the guidance block are split along y. Batch must be 1."""
        self._row_bands = {"comm": comm, "graph": bool(graph), "plans": {}}

    def disable_row_bands(self):
        rb = getattr(self, "_row_bands", None)
        if rb:
            for bg in rb["plans"].values():
                bg.close()
        self._row_bands = None

    def guidance_loss(self, latents, strength):
        """strength * wfc_loss(decode_latents_tile(latents)) as ONE fused library call, differentiable w.r.t.
        `latents`: decode, softmax, the four-direction adjacency contraction and the whole backward never leave the
        device library (the wave is not materialised in torch)."""
        rb = getattr(self, "_row_bands", None)
        if rb:
            from .rowband import BandGuidance, BandGuidanceFunction

            if latents.shape[0] != 1:
                raise ValueError("row bands split ONE map: batch size must be 1, got %d" % latents.shape[0])
            key = (latents.shape[2], latents.shape[3], str(latents.device))
            bg = rb["plans"].get(key)
            if bg is None:
                ax, ay = self.wfc_loss._mats()
                bg = BandGuidance(self.tile_vae, ax, ay, latents.shape[2], latents.shape[3], rb["comm"], latents.device)
                rb["plans"][key] = bg
            return BandGuidanceFunction.apply(latents, bg, 1.0 / LATENT_SCALE, float(strength), rb["graph"])
        t_pad = self.tile_vae.config.out_channels
        handle = self.wfc_loss.native_handle(t_pad, latents.device)
        return GuidanceLossFunction.apply(latents, self.tile_vae._native, handle, 1.0 / LATENT_SCALE, float(strength))

    def wfc_guidance(self, latents, noise_pred, t, strength):
        """The guidance block (reference :606-615): one gradient step on the latents against the WFC rule loss of the
        map the scheduler would produce from them."""
        with torch.enable_grad():
            _latents = latents.detach().requires_grad_()
            _lat_prev = self.scheduler.step(noise_pred, t, _latents).prev_sample
            _loss = self.guidance_loss(_lat_prev, strength)
            cond_grad = -torch.autograd.grad(_loss.sum(), _latents)[0]
        return latents.detach() + cond_grad.to(latents.dtype)

    def _affine_step(self, t):
        """(a, b) with scheduler.step(eps, t, x).prev_sample == a * x + b * eps, for schedulers that state them through a
        `coefficients(t)` method (DDIM eta = 0, PNDM/PLMS are affine in the sample; a maintainer adds the five lines to
        theirs); None otherwise - the generic autograd route then applies the chain factor."""
        f = getattr(self.scheduler, "coefficients", None)
        if f is None:
            return None
        a, b = f(t)
        return float(a), float(b)

    def wfc_guidance_final(self, latents, noise_pred, t, strength, n_steps):
        """The final-stage loop (reference :626-639): n_steps guidance updates re-using the last noise prediction and
        timestep. With an affine scheduler step the whole loop is ONE captured launch sequence inside the library
        (wfd_guidance_loop: scheduler step, decode, loss, backward and latent update of all n_steps iterations, no host
        round trip in between); otherwise it is n_steps calls of wfc_guidance."""
        import os

        ab = self._affine_step(t)
        fused = (ab is not None and not getattr(self, "_row_bands", None) and latents.is_cuda and cuda_graphs_enabled()
                 and latents.shape[0] <= max_micro_batch() and os.environ.get("WFD_FUSED_FINAL", "1") != "0"
                 and not torch.cuda.is_current_stream_capturing())
        if not fused:
            with self.progress_bar(total=n_steps) as bar:
                for _ in range(n_steps):
                    latents = self.wfc_guidance(latents, noise_pred, t, strength)
                    bar.update()
            return latents
        t_pad = self.tile_vae.config.out_channels
        handle = self.wfc_loss.native_handle(t_pad, latents.device)
        new, _ = guidance_loop(latents, noise_pred, self.tile_vae._native, handle, ab[0], ab[1], 1.0 / LATENT_SCALE,
                               float(strength), int(n_steps))
        return new.to(latents.dtype)

    # ------------------------------------------------------------------ third-party stages, used as handed in
    def _encode_prompt(self, prompt, device, num_images_per_prompt, do_classifier_free_guidance, negative_prompt):
        batch_size = len(prompt) if isinstance(prompt, list) else 1
        tok = self.tokenizer

        def embed(texts):
            ids = tok(texts, padding="max_length", max_length=tok.model_max_length, truncation=True, return_tensors="pt")
            use_mask = getattr(getattr(self.text_encoder, "config", None), "use_attention_mask", False)
            mask = ids.attention_mask.to(device) if use_mask else None
            e = self.text_encoder(ids.input_ids.to(device), attention_mask=mask)[0]
            n, seq, _ = e.shape
            return e.repeat(1, num_images_per_prompt, 1).view(n * num_images_per_prompt, seq, -1)

        text_embeddings = embed(prompt)
        if do_classifier_free_guidance:
            if negative_prompt is None:
                uncond = [""] * batch_size
            elif type(prompt) is not type(negative_prompt):
                raise TypeError(f"`negative_prompt` should be the same type to `prompt`, but got {type(negative_prompt)} !="
                                f" {type(prompt)}.")
            elif isinstance(negative_prompt, str):
                uncond = [negative_prompt]
            elif batch_size != len(negative_prompt):
                raise ValueError(f"`negative_prompt` has batch size {len(negative_prompt)}, but `prompt` has batch size"
                                 f" {batch_size}.")
            else:
                uncond = negative_prompt
            text_embeddings = torch.cat([embed(uncond), text_embeddings])
        return text_embeddings

    def run_safety_checker(self, image, device, dtype):
        if self.safety_checker is None:
            return image, None
        inp = self.feature_extractor(self.numpy_to_pil(image), return_tensors="pt").to(device)
        return self.safety_checker(images=image, clip_input=inp.pixel_values.to(dtype))

    def decode_latents(self, latents):
        image = self.vae.decode(1 / LATENT_SCALE * latents).sample
        image = (image / 2 + 0.5).clamp(0, 1)
        return image.cpu().permute(0, 2, 3, 1).float().numpy()

    def prepare_extra_step_kwargs(self, generator, eta):
        params = set(inspect.signature(self.scheduler.step).parameters.keys())
        kw = {}
        if "eta" in params:
            kw["eta"] = eta
        if "generator" in params:
            kw["generator"] = generator
        return kw

    def check_inputs(self, prompt, height, width, callback_steps):
        if not isinstance(prompt, str) and not isinstance(prompt, list):
            raise ValueError(f"`prompt` has to be of type `str` or `list` but is {type(prompt)}")
        if height % 8 != 0 or width % 8 != 0:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")
        if callback_steps is None or not isinstance(callback_steps, int) or callback_steps <= 0:
            raise ValueError(f"`callback_steps` has to be a positive integer but is {callback_steps} of type"
                             f" {type(callback_steps)}.")

    def prepare_latents(self, batch_size, num_channels_latents, height, width, dtype, device, generator, latents=None):
        shape = (batch_size, num_channels_latents, height // self.vae_scale_factor, width // self.vae_scale_factor)
        if isinstance(generator, list) and len(generator) != batch_size:
            raise ValueError(f"You have passed a list of generators of length {len(generator)}, but requested an"
                             f" effective batch size of {batch_size}.")
        if latents is None:
            if isinstance(generator, list):
                one = (1,) + shape[1:]
                latents = torch.cat([torch.randn(one, generator=g, device=g.device, dtype=dtype).to(device) for g in generator])
            else:
                gdev = generator.device if generator is not None else device
                latents = torch.randn(shape, generator=generator, device=gdev, dtype=dtype).to(device)
        else:
            if latents.shape != shape:
                raise ValueError(f"Unexpected latents shape, got {latents.shape}, expected {shape}")
            latents = latents.to(device)
        return latents * self.scheduler.init_noise_sigma

    # ------------------------------------------------------------------ generation
    @torch.no_grad()
    def __call__(
        self,
        prompt: Union[str, List[str]],
        height: Optional[int] = None,
        width: Optional[int] = None,
        num_inference_steps: int = 50,
        guidance_scale: float = 3.5,
        negative_prompt: Optional[Union[str, List[str]]] = None,
        num_images_per_prompt: Optional[int] = 1,
        eta: float = 0.0,
        generator: Optional[Union[torch.Generator, List[torch.Generator]]] = None,
        latents: Optional[torch.FloatTensor] = None,
        output_type: Optional[str] = "pil",
        return_dict: bool = True,
        callback: Optional[Callable[[int, int, torch.FloatTensor], None]] = None,
        callback_steps: Optional[int] = 1,
        wfc_guidance_start_step: Optional[int] = 0,
        wfc_guidance_strength: Optional[float] = 1.0,
        wfc_guidance_final_steps: Optional[int] = 0,
        wfc_guidance_final_strength: Optional[float] = 1.0,
        color_balance: Optional[List[float]] = None,
    ):
        sample_size = self.unet.config.sample_size
        height = height or sample_size * self.vae_scale_factor
        width = width or sample_size * self.vae_scale_factor
        self.check_inputs(prompt, height, width, callback_steps)
        if color_balance is not None:
            raise NotImplementedError("color_balance: the reference path for it calls an undefined function "
                                      "(pipeline_wfd.py:644) and is not part of the guidance path")
        batch_size = 1 if isinstance(prompt, str) else len(prompt)
        device = self._execution_device
        do_cfg = guidance_scale > 1.0
        text_embeddings = self._encode_prompt(prompt, device, num_images_per_prompt, do_cfg, negative_prompt)

        self.scheduler.set_timesteps(num_inference_steps, device=device)
        timesteps = self.scheduler.timesteps
        self.wfc_loss.to(device)
        latents = self.prepare_latents(batch_size * num_images_per_prompt, self.unet.in_channels, height, width,
                                       text_embeddings.dtype, device, generator, latents)
        extra_step_kwargs = self.prepare_extra_step_kwargs(generator, eta)
        latents = self._denoise(latents, timesteps, num_inference_steps, text_embeddings, do_cfg, guidance_scale,
                                extra_step_kwargs, callback, callback_steps, wfc_guidance_start_step,
                                wfc_guidance_strength, wfc_guidance_final_steps, wfc_guidance_final_strength)
        return self._finish(latents, device, text_embeddings.dtype, output_type, return_dict)

    def _denoise(self, latents, timesteps, num_inference_steps, text_embeddings, do_cfg, guidance_scale, extra_step_kwargs,
                 callback, callback_steps, wfc_guidance_start_step, wfc_guidance_strength, wfc_guidance_final_steps,
                 wfc_guidance_final_strength):
        """The denoising loop with WFC guidance (reference :590-639), shared by txt2img and img2img."""
        order = getattr(self.scheduler, "order", 1)
        num_warmup_steps = len(timesteps) - num_inference_steps * order
        noise_pred, t = None, None
        with self.progress_bar(total=num_inference_steps) as bar:
            for i, t in enumerate(timesteps):
                model_in = torch.cat([latents] * 2) if do_cfg else latents
                model_in = self.scheduler.scale_model_input(model_in, t)
                noise_pred = self.unet(model_in, t, encoder_hidden_states=text_embeddings).sample
                if do_cfg:
                    uncond, text = noise_pred.chunk(2)
                    noise_pred = uncond + guidance_scale * (text - uncond)
                if i >= wfc_guidance_start_step and wfc_guidance_strength != 0:
                    latents = self.wfc_guidance(latents, noise_pred, t, wfc_guidance_strength)
                latents = self.scheduler.step(noise_pred, t, latents, **extra_step_kwargs).prev_sample
                if i == len(timesteps) - 1 or ((i + 1) > num_warmup_steps and (i + 1) % order == 0):
                    bar.update()
                    if callback is not None and i % callback_steps == 0:
                        callback(i, t, latents)
        # final stage: guidance only, re-using the last noise prediction and timestep (reference :626-639)
        if wfc_guidance_final_steps > 0 and wfc_guidance_final_strength != 0:
            latents = self.wfc_guidance_final(latents, noise_pred, t, wfc_guidance_final_strength, wfc_guidance_final_steps)
        return latents

    def _finish(self, latents, device, dtype, output_type, return_dict):
        image = self.decode_latents(latents)
        wave = self.decode_latents_tile(latents, numpy=True)
        image, has_nsfw_concept = self.run_safety_checker(image, device, dtype)
        if output_type == "pil":
            image = self.numpy_to_pil(image)
        if not return_dict:
            return (image, has_nsfw_concept)
        return WaveFunctionDiffusionPipelineOutput(images=image, waves=wave, nsfw_content_detected=has_nsfw_concept)
