"""wavefunctiondiffusion_b200 — B200 (sm_100a) implementation of WaveFunctionDiffusion's WFC-guidance hot path behind
the reference's own Python objects (the part of `wfd/__init__.py:5-36` that is on the path):

    from wavefunctiondiffusion_b200 import WaveFunctionDiffusionPipeline, AutoencoderTile, WFCLossEinsum

Importing the package does not load the CUDA library; the first decode / loss call does, and fails loudly when
libwfd_b200.so has not been built (python -m wavefunctiondiffusion_b200.build). There is no CPU path.
"""
__version__ = "0.2.0"

from .wf_diffusers import (AutoencoderTile, WaveFunctionDiffusionImg2ImgPipeline, WaveFunctionDiffusionPipeline,
                           preprocess_img, preprocess_wave)
from .wfdf.losses import WFCLossEinsum

__all__ = ["AutoencoderTile", "WaveFunctionDiffusionPipeline", "WaveFunctionDiffusionImg2ImgPipeline", "preprocess_wave",
           "preprocess_img", "WFCLossEinsum"]
