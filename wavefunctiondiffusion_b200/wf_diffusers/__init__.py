"""The reference's `wfd.wf_diffusers` names (wfd/wf_diffusers/__init__.py:3-5), served by libwfd_b200.so:

    from wavefunctiondiffusion_b200.wf_diffusers import WaveFunctionDiffusionPipeline, AutoencoderTile
"""
from .wf_vae import AutoencoderTile
from .pipeline_wfd import WaveFunctionDiffusionPipeline, WaveFunctionDiffusionPipelineOutput
from .pipeline_wfd_img2img import WaveFunctionDiffusionImg2ImgPipeline, preprocess_wave, preprocess as preprocess_img

__all__ = ["AutoencoderTile", "WaveFunctionDiffusionPipeline", "WaveFunctionDiffusionPipelineOutput",
           "WaveFunctionDiffusionImg2ImgPipeline", "preprocess_wave", "preprocess_img"]
