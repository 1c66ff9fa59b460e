"""B200-native LTX-2 DiT denoise-step forward (drop-in for mlx_video.models.ltx.LTXModel).

The directory is named ``mlx-video_b200`` (not importable as such); import it as ``mlx_video_b200``
through the loader module of that name at the repository root.  Importing the package loads
``csrc/libltxb.so`` (the C ABI of include/ltxb.h) and fails if it is missing: there is no CPU or
eager fallback on this path.
"""
__version__ = "0.1.0"

from . import _lib, ops  # noqa: F401  (loads libltxb.so)
from ._lib import LtxbError  # noqa: F401
from .config import (AttentionType, LTXModelConfig, LTXModelType, LTXRopeType, TransformerConfig,  # noqa: F401
                     production_config)
from .lora import LoraSpec, apply_lora_to_model, apply_lora_to_weights  # noqa: F401
from .model import AdaLayerNormSingle, LTXModel, PixArtAlphaTextProjection, X0Model, to_denoised  # noqa: F401
from .rope import precompute_freqs_cis  # noqa: F401
from .upsampler import LatentUpsampler, load_upsampler, upsample_latents  # noqa: F401
from .transformer import (Attention, BasicAVTransformerBlock, FeedForward, Modality, TransformerArgs,  # noqa: F401
                          Workspace)
