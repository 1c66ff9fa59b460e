"""Configuration records of the LTX-2 DiT, field-for-field compatible with the reference's
``mlx_video/models/ltx/config.py:7-23,56-61,93-182`` so that a caller's ``LTXModelConfig(...)``
(production values: ``mlx_video/generate.py:2866-2891``) constructs this model unchanged.
"""
from __future__ import annotations

import inspect
from dataclasses import dataclass, fields
from enum import Enum
from typing import Any, Dict, List, Optional


class LTXModelType(Enum):
    AudioVideo = "ltx av model"
    VideoOnly = "ltx video only model"
    AudioOnly = "ltx audio only model"

    def is_video_enabled(self) -> bool:
        return self is not LTXModelType.AudioOnly

    def is_audio_enabled(self) -> bool:
        return self is not LTXModelType.VideoOnly


class LTXRopeType(Enum):
    INTERLEAVED = "interleaved"
    SPLIT = "split"
    TWO_D = "2d"


class AttentionType(Enum):
    DEFAULT = "default"


class _DictMixin:
    @classmethod
    def from_dict(cls, params: Dict[str, Any]):
        accepted = inspect.signature(cls).parameters
        return cls(**{k: v for k, v in params.items() if k in accepted})

    def to_dict(self) -> Dict[str, Any]:
        out: Dict[str, Any] = {}
        for f in fields(self):
            v = getattr(self, f.name)
            if v is None:
                continue
            out[f.name] = v.value if isinstance(v, Enum) else (v.to_dict() if hasattr(v, "to_dict") else v)
        return out


@dataclass
class TransformerConfig(_DictMixin):
    """Per-modality width record handed to each block (config.py:56-61)."""

    dim: int
    heads: int
    d_head: int
    context_dim: int


@dataclass
class LTXModelConfig(_DictMixin):
    model_type: LTXModelType = LTXModelType.AudioVideo

    num_attention_heads: int = 32
    attention_head_dim: int = 128
    in_channels: int = 128
    out_channels: int = 128
    num_layers: int = 48
    cross_attention_dim: int = 4096
    caption_channels: int = 3840

    audio_num_attention_heads: int = 32
    audio_attention_head_dim: int = 64
    audio_in_channels: int = 128
    audio_out_channels: int = 128
    audio_cross_attention_dim: int = 2048
    audio_caption_channels: int = 3840

    positional_embedding_theta: float = 10000.0
    positional_embedding_max_pos: Optional[List[int]] = None
    audio_positional_embedding_max_pos: Optional[List[int]] = None
    use_middle_indices_grid: bool = True
    rope_type: LTXRopeType = LTXRopeType.INTERLEAVED
    double_precision_rope: bool = False

    timestep_scale_multiplier: int = 1000
    av_ca_timestep_scale_multiplier: int = 1000

    norm_eps: float = 1e-6
    attention_type: AttentionType = AttentionType.DEFAULT
    vae_config: Optional[Any] = None  # carried for from_dict round trips; the VAE is outside this path

    def __post_init__(self) -> None:
        if self.positional_embedding_max_pos is None:
            self.positional_embedding_max_pos = [20, 2048, 2048]
        if self.audio_positional_embedding_max_pos is None:
            self.audio_positional_embedding_max_pos = [20]
        if isinstance(self.model_type, str):
            self.model_type = LTXModelType(self.model_type)
        if isinstance(self.rope_type, str):
            self.rope_type = LTXRopeType(self.rope_type)
        if isinstance(self.attention_type, str):
            self.attention_type = AttentionType(self.attention_type)

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim

    @property
    def audio_inner_dim(self) -> int:
        return self.audio_num_attention_heads * self.audio_attention_head_dim

    def get_video_config(self) -> Optional[TransformerConfig]:
        if not self.model_type.is_video_enabled():
            return None
        return TransformerConfig(self.inner_dim, self.num_attention_heads, self.attention_head_dim,
                                 self.cross_attention_dim)

    def get_audio_config(self) -> Optional[TransformerConfig]:
        if not self.model_type.is_audio_enabled():
            return None
        return TransformerConfig(self.audio_inner_dim, self.audio_num_attention_heads, self.audio_attention_head_dim,
                                 self.audio_cross_attention_dim)


def production_config(model_type: LTXModelType = LTXModelType.VideoOnly, num_layers: int = 48) -> LTXModelConfig:
    """The configuration every real entry point of the reference builds (generate.py:2866-2891)."""
    return LTXModelConfig(
        model_type=model_type, num_attention_heads=32, attention_head_dim=128, in_channels=128, out_channels=128,
        num_layers=num_layers, cross_attention_dim=4096, caption_channels=3840, audio_num_attention_heads=32,
        audio_attention_head_dim=64, audio_in_channels=128, audio_out_channels=128, audio_cross_attention_dim=2048,
        audio_caption_channels=3840, positional_embedding_theta=10000.0, positional_embedding_max_pos=[20, 2048, 2048],
        audio_positional_embedding_max_pos=[20], use_middle_indices_grid=True, rope_type=LTXRopeType.SPLIT,
        double_precision_rope=True, timestep_scale_multiplier=1000, av_ca_timestep_scale_multiplier=1000, norm_eps=1e-6)
