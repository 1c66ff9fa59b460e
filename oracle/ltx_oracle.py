"""ORACLE — test infrastructure, not product code.

CPU restatement (PyTorch fp32, optionally fp64) of the reference's LTX-2 DiT forward, i.e. of the
MLX code under /root/reference/mlx_video/models/ltx/ plus the sampler-side helpers in
mlx_video/generate.py and mlx_video/utils.py.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module; the product
package (mlx-video_b200/) never does.

Parity pin: the reference ships no golden vectors for this path (tests/test_heavy_pipeline_parity.py:33,53
assert only that a file exists) and its runtime dependency ``mlx==0.30.1`` (uv.lock:760-761; Apple-only,
absent from /root/reference and from this image) cannot run here.  The oracle is therefore pinned
against the reference's OWN Python sources executed over ``oracle/mlx_shim`` (an mlx.core / mlx.nn
stand-in backed by torch): ``oracle/make_golden.py`` imports /root/reference's transformer.py,
attention.py, rope.py, adaln.py, feed_forward.py, text_projection.py, ltx.py and the grid / scheduler
functions of generate.py unmodified, runs them, and commits the outputs under tests/golden/.  What is
restated rather than executed is exactly the third-party MLX library arithmetic (nn.Linear,
mx.fast.rms_norm, mx.fast.scaled_dot_product_attention, nn.gelu_approx, nn.SiLU, nn.LayerNorm,
cos/sin/power/linspace), with their published definitions.

Every function cites the reference file:line it follows (paths relative to /root/reference/).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field, replace
from enum import Enum
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------------------
# config  (mlx_video/models/ltx/config.py:7-23,56-61,93-182; production values generate.py:2866-2891)
# --------------------------------------------------------------------------------------------------
class LTXModelType(Enum):
    AudioVideo = "ltx av model"
    VideoOnly = "ltx video only model"
    AudioOnly = "ltx audio only model"

    def is_video_enabled(self) -> bool:
        return self in (LTXModelType.AudioVideo, LTXModelType.VideoOnly)

    def is_audio_enabled(self) -> bool:
        return self in (LTXModelType.AudioVideo, LTXModelType.AudioOnly)


class LTXRopeType(Enum):
    INTERLEAVED = "interleaved"
    SPLIT = "split"


@dataclass
class OracleConfig:
    model_type: LTXModelType = LTXModelType.VideoOnly
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
    positional_embedding_max_pos: List[int] = field(default_factory=lambda: [20, 2048, 2048])
    audio_positional_embedding_max_pos: List[int] = field(default_factory=lambda: [20])
    use_middle_indices_grid: bool = True
    rope_type: LTXRopeType = LTXRopeType.SPLIT
    double_precision_rope: bool = True
    timestep_scale_multiplier: int = 1000
    av_ca_timestep_scale_multiplier: int = 1000
    norm_eps: float = 1e-6

    @property
    def inner_dim(self) -> int:  # config.py:149-152
        return self.num_attention_heads * self.attention_head_dim

    @property
    def audio_inner_dim(self) -> int:  # config.py:154-157
        return self.audio_num_attention_heads * self.audio_attention_head_dim


@dataclass(frozen=True)
class Modality:  # transformer.py:13-22
    latent: Tensor
    timesteps: Tensor
    positions: Tensor
    context: Tensor
    enabled: bool = True
    context_mask: Optional[Tensor] = None
    positional_embeddings: Optional[Tuple[Tensor, Tensor]] = None


@dataclass(frozen=True)
class TransformerArgs:  # transformer.py:25-36
    x: Tensor
    context: Tensor
    context_mask: Optional[Tensor]
    timesteps: Tensor
    embedded_timestep: Tensor
    positional_embeddings: Tuple[Tensor, Tensor]
    cross_positional_embeddings: Optional[Tuple[Tensor, Tensor]]
    cross_scale_shift_timestep: Optional[Tensor]
    cross_gate_timestep: Optional[Tensor]
    enabled: bool


# --------------------------------------------------------------------------------------------------
# sampler-side helpers (numpy, as the reference)
# --------------------------------------------------------------------------------------------------
STAGE_1_SIGMAS = [1.0, 0.99375, 0.9875, 0.98125, 0.975, 0.909375, 0.725, 0.421875, 0.0]  # generate.py:339
STAGE_2_SIGMAS = [0.909375, 0.725, 0.421875, 0.0]  # generate.py:340
BASE_SHIFT_ANCHOR = 1024  # generate.py:343
MAX_SHIFT_ANCHOR = 4096  # generate.py:344
AUDIO_LATENT_SAMPLE_RATE = 16000  # generate.py:348
AUDIO_HOP_LENGTH = 160  # generate.py:349
AUDIO_LATENT_DOWNSAMPLE_FACTOR = 4  # generate.py:350
AUDIO_LATENTS_PER_SECOND = AUDIO_LATENT_SAMPLE_RATE / AUDIO_HOP_LENGTH / AUDIO_LATENT_DOWNSAMPLE_FACTOR  # :353


def create_position_grid(batch_size, num_frames, height, width, temporal_scale=8, spatial_scale=32, fps=24.0,
                         causal_fix=True) -> np.ndarray:
    """generate.py:470-525 — (B,3,T,2) fp32 [start,end) bounds, token order f*H*W + h*W + w."""
    t = np.arange(0, num_frames, 1)
    h = np.arange(0, height, 1)
    w = np.arange(0, width, 1)
    tg, hg, wg = np.meshgrid(t, h, w, indexing="ij")
    starts = np.stack([tg, hg, wg], axis=0)
    ends = starts + np.array([1, 1, 1]).reshape(3, 1, 1, 1)
    coords = np.stack([starts, ends], axis=-1).reshape(3, num_frames * height * width, 2)
    coords = np.tile(coords[np.newaxis, ...], (batch_size, 1, 1, 1))
    scale = np.array([temporal_scale, spatial_scale, spatial_scale]).reshape(1, 3, 1, 1)
    px = (coords * scale).astype(np.float32)
    if causal_fix:
        px[:, 0, :, :] = np.clip(px[:, 0, :, :] + 1 - temporal_scale, a_min=0, a_max=None)
    px[:, 0, :, :] = px[:, 0, :, :] / fps
    return px.astype(np.float32)


def create_audio_position_grid(batch_size, audio_frames, sample_rate=AUDIO_LATENT_SAMPLE_RATE,
                               hop_length=AUDIO_HOP_LENGTH, downsample_factor=AUDIO_LATENT_DOWNSAMPLE_FACTOR,
                               is_causal=True) -> np.ndarray:
    """generate.py:528-551 — (B,1,Ta,2) fp32 seconds."""

    def sec(a, b):
        lf = np.arange(a, b, dtype=np.float32)
        mel = lf * downsample_factor
        if is_causal:
            mel = np.clip(mel + 1 - downsample_factor, 0, None)
        return mel * hop_length / sample_rate

    pos = np.stack([sec(0, audio_frames), sec(1, audio_frames + 1)], axis=-1)[np.newaxis, np.newaxis]
    return np.tile(pos, (batch_size, 1, 1, 1)).astype(np.float32)


def compute_audio_frames(num_video_frames: int, fps: float) -> int:
    """generate.py:554-557"""
    return round(num_video_frames / fps * AUDIO_LATENTS_PER_SECOND)


def ltx2_scheduler(steps, num_tokens=None, max_shift=2.05, base_shift=0.95, stretch=True, terminal=0.1) -> np.ndarray:
    """generate.py:410-467 — fp32 sigma schedule of length steps+1."""
    tokens = MAX_SHIFT_ANCHOR if num_tokens is None else min(num_tokens, MAX_SHIFT_ANCHOR)
    sigmas = np.linspace(1.0, 0.0, steps + 1)
    mm = (max_shift - base_shift) / (MAX_SHIFT_ANCHOR - BASE_SHIFT_ANCHOR)
    b = base_shift - mm * BASE_SHIFT_ANCHOR
    shift = tokens * mm + b
    out = np.zeros_like(sigmas)
    nz = sigmas != 0
    if np.any(nz):
        out[nz] = math.exp(shift) / (math.exp(shift) + (1 / sigmas[nz] - 1) ** 1)
    sigmas = out
    if stretch:
        nzm = sigmas != 0
        one_minus = 1.0 - sigmas[nzm]
        sf = one_minus[-1] / (1.0 - terminal)
        if np.isfinite(sf) and sf != 0:
            sigmas[nzm] = 1.0 - (one_minus / sf)
    return sigmas.astype(np.float32)


def to_denoised(noisy: Tensor, velocity: Tensor, sigma) -> Tensor:
    """utils.py:404-440 — x0 = x - sigma * v in fp32, sigma right-padded with unit axes."""
    n32, v32 = noisy.float(), velocity.float()
    if isinstance(sigma, (int, float)):
        s32 = torch.tensor(sigma, dtype=torch.float32)
    else:
        s32 = sigma.float()
        while s32.dim() < v32.dim():
            s32 = s32.unsqueeze(-1)
    return (n32 - s32 * v32).to(noisy.dtype)


def cfg_combine(v_pos: Tensor, v_neg: Tensor, scale: float) -> Tensor:
    """generate.py:1255,1283 — v_pos + (s-1)(v_pos - v_neg)."""
    return v_pos + (scale - 1.0) * (v_pos - v_neg)


def euler_step(latents: Tensor, denoised: Tensor, sigma: float, sigma_next: float) -> Tensor:
    """generate.py:1293-1301 (fp32_euler=True): x0 + sigma_next * (x - x0) / sigma."""
    l32, d32 = latents.float(), denoised.float()
    s, sn = torch.tensor(sigma, dtype=torch.float32), torch.tensor(sigma_next, dtype=torch.float32)
    return (d32 + sn * (l32 - d32) / s).to(latents.dtype)


def apply_denoise_mask(denoised: Tensor, clean: Tensor, mask: Tensor) -> Tensor:
    """conditioning/latent.py:180-196 — denoised*mask + clean*(1-mask)."""
    return denoised * mask + clean * (1 - mask)


# --------------------------------------------------------------------------------------------------
# RoPE  (rope.py)
# --------------------------------------------------------------------------------------------------
def rope_freq_indices(theta: float, n_pos_dims: int, dim: int) -> Tensor:
    """rope.py:446-457 — theta ** linspace(0,1,dim//(2*n_pos_dims)) * pi/2, fp32."""
    num = dim // (2 * n_pos_dims)
    if num == 0:
        num = 1
    log_start = math.log(1.0) / math.log(theta)
    log_end = math.log(theta) / math.log(theta)
    lin = torch.linspace(log_start, log_end, num, dtype=torch.float32)
    return torch.pow(torch.tensor(theta, dtype=torch.float32), lin) * (math.pi / 2)


def precompute_freqs_cis(indices_grid: Tensor, dim: int, theta: float = 10000.0, max_pos: Optional[List[int]] = None,
                         use_middle_indices_grid: bool = False, num_attention_heads: int = 32,
                         rope_type: LTXRopeType = LTXRopeType.INTERLEAVED, double_precision: bool = False):
    """rope.py:364-416 (dispatch) and :419-529 (the fp32 "double precision" path; the plain path :399-414
    computes the same quantities through generate_freq_grid/generate_freqs :175-291)."""
    if max_pos is None:
        max_pos = [20, 2048, 2048]
    grid = indices_grid.float()
    n_pos = grid.shape[1]
    n_elem = 2 * n_pos
    freq = rope_freq_indices(theta, n_pos, dim)
    if use_middle_indices_grid:  # rope.py:461-466
        assert grid.dim() == 4 and grid.shape[-1] == 2
        grid = (grid[..., 0] + grid[..., 1]) / 2.0
    elif grid.dim() == 4:
        grid = grid[..., 0]
    assert n_pos == len(max_pos), "Number of position dimensions must match max_pos length"  # rope.py:228
    frac = torch.stack([grid[:, i, :] / max_pos[i] for i in range(n_pos)], dim=-1)  # (B,T,n) :473-480
    scaled = frac * 2 - 1  # :483
    freqs = scaled.unsqueeze(-1) * freq.reshape(1, 1, 1, -1)  # (B,T,n,F) :488
    freqs = freqs.transpose(-1, -2).reshape(freqs.shape[0], freqs.shape[1], -1)  # (B,T,F*n) :491-493
    cos, sin = torch.cos(freqs), torch.sin(freqs)
    if rope_type == LTXRopeType.SPLIT:  # :499-516
        pad = dim // 2 - cos.shape[-1]
        if pad > 0:
            cos = torch.cat([torch.ones(*cos.shape[:-1], pad), cos], dim=-1)
            sin = torch.cat([torch.zeros(*sin.shape[:-1], pad), sin], dim=-1)
        b, t = cos.shape[0], cos.shape[1]
        cos = cos.reshape(b, t, num_attention_heads, -1).transpose(1, 2)
        sin = sin.reshape(b, t, num_attention_heads, -1).transpose(1, 2)
    else:  # :517-527
        cos = cos.repeat_interleave(2, dim=-1)
        sin = sin.repeat_interleave(2, dim=-1)
        pad = dim % n_elem
        if pad > 0:
            cos = torch.cat([torch.ones(*cos.shape[:-1], pad), cos], dim=-1)
            sin = torch.cat([torch.zeros(*sin.shape[:-1], pad), sin], dim=-1)
    return cos.contiguous(), sin.contiguous()


def apply_split_rotary_emb(x: Tensor, cos: Tensor, sin: Tensor) -> Tensor:
    """rope.py:109-172 — per head, halves [a;b] -> [a cos - b sin ; b cos + a sin], fp32."""
    dtype = x.dtype
    needs = False
    if x.dim() != 4 and cos.dim() == 4:
        b, h, t, _ = cos.shape
        x = x.reshape(b, t, h, -1).transpose(1, 2)
        needs = True
    x = x.float()
    d = x.shape[-1]
    sp = x.reshape(*x.shape[:-1], 2, d // 2)
    first, second = sp[..., 0, :], sp[..., 1, :]
    of = first * cos - sin * second
    os_ = second * cos + sin * first
    out = torch.stack([of, os_], dim=-2).reshape(x.shape)
    if needs:
        b, h, t, d = out.shape
        out = out.transpose(1, 2).reshape(b, t, h * d)
    return out.to(dtype)


def apply_interleaved_rotary_emb(x: Tensor, cos: Tensor, sin: Tensor) -> Tensor:
    """rope.py:33-75 — adjacent pairs (x0,x1) -> x*cos + (-x1,x0)*sin."""
    dtype = x.dtype
    x = x.float()
    shp = x.shape
    xp = x.reshape(*shp[:-1], shp[-1] // 2, 2)
    rot = torch.stack([-xp[..., 1], xp[..., 0]], dim=-1).reshape(shp)
    return (x * cos.float() + rot * sin.float()).to(dtype)


def apply_rotary_emb(x, freqs_cis, rope_type):  # rope.py:9-30
    if rope_type == LTXRopeType.INTERLEAVED:
        return apply_interleaved_rotary_emb(x, freqs_cis[0], freqs_cis[1])
    if rope_type == LTXRopeType.SPLIT:
        return apply_split_rotary_emb(x, freqs_cis[0], freqs_cis[1])
    raise ValueError(f"Invalid rope type: {rope_type}")


# --------------------------------------------------------------------------------------------------
# ops
# --------------------------------------------------------------------------------------------------
def rms_norm(x: Tensor, eps: float = 1e-6) -> Tensor:
    """utils.py:398-400 — mx.fast.rms_norm with unit weight: x * rsqrt(mean(x^2) + eps)."""
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)


def get_timestep_embedding(timesteps: Tensor, dim: int, flip_sin_to_cos=False, downscale_freq_shift=1.0, scale=1.0,
                           max_period=10000) -> Tensor:
    """utils.py:486-526"""
    assert timesteps.dim() == 1
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(0, half, dtype=torch.float32)
    exponent = exponent / (half - downscale_freq_shift)
    emb = torch.exp(exponent)
    emb = (timesteps[:, None].float() * scale) * emb[None, :]
    emb = torch.cat([torch.cos(emb), torch.sin(emb)], -1) if flip_sin_to_cos else torch.cat([torch.sin(emb), torch.cos(emb)], -1)
    if dim % 2 == 1:
        emb = F.pad(emb, (0, 1))
    return emb


class Params:
    """Flat name -> tensor store using the reference's sanitized state-dict names (ltx.py:508-533)."""

    def __init__(self, tensors: Dict[str, Tensor], prefix: str = ""):
        self.t = tensors
        self.prefix = prefix

    def sub(self, name: str) -> "Params":
        return Params(self.t, f"{self.prefix}{name}.")

    def __getitem__(self, name: str) -> Tensor:
        return self.t[self.prefix + name]

    def has(self, name: str) -> bool:
        return (self.prefix + name) in self.t


def linear(p: Params, x: Tensor) -> Tensor:
    """mlx nn.Linear: x W^T + b, W stored (out,in)."""
    return F.linear(x, p["weight"], p["bias"] if p.has("bias") else None)


def adaln_single(p: Params, timestep: Tensor) -> Tuple[Tensor, Tensor]:
    """adaln.py:29-47 with :66-67,81-85,134-138: sinusoid(256, flip, shift 0) -> Linear -> SiLU -> Linear
    (= embedded_timestep) -> SiLU -> Linear (k*D)."""
    proj = get_timestep_embedding(timestep, 256, flip_sin_to_cos=True, downscale_freq_shift=0)
    te = p.sub("emb").sub("timestep_embedder")
    e = linear(te.sub("linear2"), F.silu(linear(te.sub("linear1"), proj)))
    return linear(p.sub("linear"), F.silu(e)), e


def text_projection(p: Params, x: Tensor) -> Tensor:
    """text_projection.py:22-26"""
    return linear(p.sub("linear2"), F.gelu(linear(p.sub("linear1"), x), approximate="tanh"))


def feed_forward(p: Params, x: Tensor) -> Tensor:
    """feed_forward.py:35-40"""
    return linear(p.sub("proj_out"), F.gelu(linear(p.sub("proj_in"), x), approximate="tanh"))


def scaled_dot_product_attention(q: Tensor, k: Tensor, v: Tensor, heads: int, mask: Optional[Tensor] = None) -> Tensor:
    """attention.py:13-53 — (B,T,H*dh) -> heads, softmax(QK^T/sqrt(dh) + mask) V, back."""
    b, tq, dim = q.shape
    tk = k.shape[1]
    dh = dim // heads
    q = q.reshape(b, tq, heads, dh).transpose(1, 2)
    k = k.reshape(b, tk, heads, dh).transpose(1, 2)
    v = v.reshape(b, tk, heads, dh).transpose(1, 2)
    if mask is not None:
        if mask.dim() == 2:
            mask = mask.unsqueeze(0)
        if mask.dim() == 3:
            mask = mask.unsqueeze(1)
    s = (q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(dh))
    if mask is not None:
        s = s + mask
    out = torch.softmax(s.float(), dim=-1).to(v.dtype) @ v
    return out.transpose(1, 2).reshape(b, tq, heads * dh)


def attention(p: Params, x: Tensor, heads: int, rope_type: LTXRopeType, eps: float, context: Optional[Tensor] = None,
              mask: Optional[Tensor] = None, pe=None, k_pe=None) -> Tensor:
    """attention.py:102-142"""
    q = linear(p.sub("to_q"), x)
    context = x if context is None else context
    k = linear(p.sub("to_k"), context)
    v = linear(p.sub("to_v"), context)
    q = rms_norm(q, eps) * p["q_norm.weight"]  # nn.RMSNorm over the full inner dim (attention.py:96-97)
    k = rms_norm(k, eps) * p["k_norm.weight"]
    if pe is not None:
        q = apply_rotary_emb(q, pe, rope_type)
        k = apply_rotary_emb(k, pe if k_pe is None else k_pe, rope_type)
    return linear(p.sub("to_out"), scaled_dot_product_attention(q, k, v, heads, mask))


def get_ada_values(table: Tensor, batch_size: int, timestep: Tensor, indices: slice):
    """transformer.py:135-177"""
    n = table.shape[0]
    ts = timestep.reshape(batch_size, timestep.shape[1], n, -1)[:, :, indices, :]
    ada = table[indices][None, None] + ts
    return tuple(ada[:, :, i, :] for i in range(ada.shape[2]))


def get_av_ca_ada_values(table: Tensor, batch_size: int, scale_shift_ts: Tensor, gate_ts: Tensor, n_ss: int = 4):
    """transformer.py:179-219"""
    ss = get_ada_values(table[:n_ss], batch_size, scale_shift_ts, slice(None, None))
    g = get_ada_values(table[n_ss:], batch_size, gate_ts, slice(None, None))
    sq = lambda t: t.squeeze(1) if t.shape[1] == 1 else t  # noqa: E731
    return (*[sq(t) for t in ss], *[sq(t) for t in g])


def transformer_block(p: Params, cfg: OracleConfig, video: Optional[TransformerArgs], audio: Optional[TransformerArgs]):
    """transformer.py:221-361"""
    eps, rt = cfg.norm_eps, cfg.rope_type
    vh, ah = cfg.num_attention_heads, cfg.audio_num_attention_heads
    vx = video.x if video is not None else None
    ax = audio.x if audio is not None else None
    run_vx = video is not None and video.enabled and vx.numel() > 0
    run_ax = audio is not None and audio.enabled and ax.numel() > 0
    run_a2v = run_vx and run_ax
    run_v2a = run_ax and run_vx
    if run_vx:
        sh, sc, g = get_ada_values(p["scale_shift_table"], vx.shape[0], video.timesteps, slice(0, 3))
        nvx = rms_norm(vx, eps) * (1 + sc) + sh
        vx = vx + attention(p.sub("attn1"), nvx, vh, rt, eps, pe=video.positional_embeddings) * g
        vx = vx + attention(p.sub("attn2"), rms_norm(vx, eps), vh, rt, eps, context=video.context, mask=video.context_mask)
    if run_ax:
        sh, sc, g = get_ada_values(p["audio_scale_shift_table"], ax.shape[0], audio.timesteps, slice(0, 3))
        nax = rms_norm(ax, eps) * (1 + sc) + sh
        ax = ax + attention(p.sub("audio_attn1"), nax, ah, rt, eps, pe=audio.positional_embeddings) * g
        ax = ax + attention(p.sub("audio_attn2"), rms_norm(ax, eps), ah, rt, eps, context=audio.context, mask=audio.context_mask)
    if run_a2v or run_v2a:
        vn3, an3 = rms_norm(vx, eps), rms_norm(ax, eps)
        sc_a_a2v, sh_a_a2v, sc_a_v2a, sh_a_v2a, gate_v2a = get_av_ca_ada_values(
            p["scale_shift_table_a2v_ca_audio"], ax.shape[0], audio.cross_scale_shift_timestep, audio.cross_gate_timestep)
        sc_v_a2v, sh_v_a2v, sc_v_v2a, sh_v_v2a, gate_a2v = get_av_ca_ada_values(
            p["scale_shift_table_a2v_ca_video"], vx.shape[0], video.cross_scale_shift_timestep, video.cross_gate_timestep)
        if run_a2v:
            vs = vn3 * (1 + sc_v_a2v) + sh_v_a2v
            as_ = an3 * (1 + sc_a_a2v) + sh_a_a2v
            vx = vx + attention(p.sub("audio_to_video_attn"), vs, ah, rt, eps, context=as_,
                                pe=video.cross_positional_embeddings, k_pe=audio.cross_positional_embeddings) * gate_a2v
        if run_v2a:
            as_ = an3 * (1 + sc_a_v2a) + sh_a_v2a
            vs = vn3 * (1 + sc_v_v2a) + sh_v_v2a
            ax = ax + attention(p.sub("video_to_audio_attn"), as_, ah, rt, eps, context=vs,
                                pe=audio.cross_positional_embeddings, k_pe=video.cross_positional_embeddings) * gate_v2a
    if run_vx:
        sh, sc, g = get_ada_values(p["scale_shift_table"], vx.shape[0], video.timesteps, slice(3, None))
        vx = vx + feed_forward(p.sub("ff"), rms_norm(vx, eps) * (1 + sc) + sh) * g
    if run_ax:
        sh, sc, g = get_ada_values(p["audio_scale_shift_table"], ax.shape[0], audio.timesteps, slice(3, None))
        ax = ax + feed_forward(p.sub("audio_ff"), rms_norm(ax, eps) * (1 + sc) + sh) * g
    return (replace(video, x=vx) if video is not None else None, replace(audio, x=ax) if audio is not None else None)


# --------------------------------------------------------------------------------------------------
# model  (ltx.py:33-506, 888-906)
# --------------------------------------------------------------------------------------------------
class OracleLTXModel:
    def __init__(self, config: OracleConfig, tensors: Dict[str, Tensor]):
        self.config = config
        self.p = Params(tensors)
        self.model_type = config.model_type

    # -- preprocessors: ltx.py:61-158 (simple), :201-247 (multi-modal)
    def _prepare(self, m: Modality, audio: bool) -> TransformerArgs:
        c, p = self.config, self.p
        pre = "audio_" if audio else ""
        inner = c.audio_inner_dim if audio else c.inner_dim
        heads = c.audio_num_attention_heads if audio else c.num_attention_heads
        max_pos = c.audio_positional_embedding_max_pos if audio else c.positional_embedding_max_pos
        x = linear(p.sub(pre + "patchify_proj"), m.latent)  # ltx.py:130
        B = x.shape[0]
        ts = m.timesteps * c.timestep_scale_multiplier  # ltx.py:68
        emb, e = adaln_single(p.sub(pre + "adaln_single"), ts.reshape(-1))
        emb = emb.reshape(B, -1, emb.shape[-1])
        e = e.reshape(B, -1, e.shape[-1])
        ctx = text_projection(p.sub(pre + "caption_projection"), m.context).reshape(B, -1, x.shape[-1])  # ltx.py:77-89
        mask = m.context_mask
        if mask is not None and not mask.dtype.is_floating_point:  # ltx.py:91-107
            mask = ((mask.to(m.latent.dtype) - 1) * 1e9).reshape(mask.shape[0], 1, -1, mask.shape[-1])
        pe = m.positional_embeddings
        if pe is None:  # ltx.py:136-145
            pe = precompute_freqs_cis(m.positions, inner, c.positional_embedding_theta, max_pos, c.use_middle_indices_grid,
                                      heads, c.rope_type, c.double_precision_rope)
        args = TransformerArgs(x, ctx, mask, emb, e, pe, None, None, None, m.enabled)
        if c.model_type == LTXModelType.AudioVideo:  # ltx.py:201-247
            cross_max = max(c.positional_embedding_max_pos[0], c.audio_positional_embedding_max_pos[0])  # :279-282
            cross_pe = precompute_freqs_cis(m.positions[:, 0:1, :], c.audio_cross_attention_dim, c.positional_embedding_theta,
                                            [cross_max], True, heads, c.rope_type, c.double_precision_rope)
            ss_name = "av_ca_audio_scale_shift_adaln_single" if audio else "av_ca_video_scale_shift_adaln_single"
            g_name = "av_ca_v2a_gate_adaln_single" if audio else "av_ca_a2v_gate_adaln_single"
            factor = c.av_ca_timestep_scale_multiplier / c.timestep_scale_multiplier
            ss, _ = adaln_single(p.sub(ss_name), ts.reshape(-1))
            gt, _ = adaln_single(p.sub(g_name), ts.reshape(-1) * factor)
            args = replace(args, cross_positional_embeddings=cross_pe,
                           cross_scale_shift_timestep=ss.reshape(B, -1, ss.shape[-1]),
                           cross_gate_timestep=gt.reshape(B, -1, gt.shape[-1]))
        return args

    def _output(self, table: Tensor, proj: Params, x: Tensor, e: Tensor) -> Tensor:
        """ltx.py:432-457 — LayerNorm(no affine) * (1+scale) + shift, shift=row 0, scale=row 1."""
        ss = table[None, None, :, :] + e[:, :, None, :]
        shift, scale = ss[:, :, 0, :], ss[:, :, 1, :]
        x = F.layer_norm(x, (x.shape[-1],), eps=self.config.norm_eps)
        return linear(proj, x * (1 + scale) + shift)

    def prepare(self, video: Optional[Modality], audio: Optional[Modality]):
        if not self.model_type.is_video_enabled() and video is not None:
            raise ValueError("Video is not enabled for this model")  # ltx.py:466-467
        if not self.model_type.is_audio_enabled() and audio is not None:
            raise ValueError("Audio is not enabled for this model")  # ltx.py:468-469
        va = self._prepare(video, False) if video is not None else None
        aa = self._prepare(audio, True) if audio is not None else None
        return va, aa

    def block(self, idx: int, va, aa):
        return transformer_block(self.p.sub(f"transformer_blocks.{idx}"), self.config, va, aa)

    def __call__(self, video: Optional[Modality] = None, audio: Optional[Modality] = None):
        """ltx.py:459-506"""
        va, aa = self.prepare(video, audio)
        for i in range(self.config.num_layers):  # ltx.py:428-429
            va, aa = self.block(i, va, aa)
        vx = self._output(self.p["scale_shift_table"], self.p.sub("proj_out"), va.x, va.embedded_timestep) if va is not None else None
        ax = self._output(self.p["audio_scale_shift_table"], self.p.sub("audio_proj_out"), aa.x, aa.embedded_timestep) if aa is not None else None
        return vx, ax


def x0_model(model: OracleLTXModel, video: Optional[Modality] = None, audio: Optional[Modality] = None):
    """ltx.py:888-906"""
    vx, ax = model(video, audio)
    dv = to_denoised(video.latent, vx, video.timesteps) if vx is not None else None
    da = to_denoised(audio.latent, ax, audio.timesteps) if ax is not None else None
    return dv, da


# --------------------------------------------------------------------------------------------------
# random-init scheme (the reference has none: tables are zeros, Linear = MLX default U(+-1/sqrt(in)))
# --------------------------------------------------------------------------------------------------
def _lin(out: Dict[str, Tensor], name: str, n_out: int, n_in: int, g: torch.Generator) -> None:
    k = 1.0 / math.sqrt(n_in)
    out[name + ".weight"] = (torch.rand(n_out, n_in, generator=g) * 2 - 1) * k
    out[name + ".bias"] = (torch.rand(n_out, generator=g) * 2 - 1) * k


def _attn(out, name, q_dim, ctx_dim, inner, g):
    _lin(out, name + ".to_q", inner, q_dim, g)
    _lin(out, name + ".to_k", inner, ctx_dim, g)
    _lin(out, name + ".to_v", inner, ctx_dim, g)
    _lin(out, name + ".to_out", q_dim, inner, g)
    out[name + ".q_norm.weight"] = 1 + 0.1 * torch.randn(inner, generator=g)
    out[name + ".k_norm.weight"] = 1 + 0.1 * torch.randn(inner, generator=g)


def _adaln(out, name, dim, coeff, g):
    _lin(out, name + ".emb.timestep_embedder.linear1", dim, 256, g)
    _lin(out, name + ".emb.timestep_embedder.linear2", dim, dim, g)
    _lin(out, name + ".linear", coeff * dim, dim, g)


def init_params(cfg: OracleConfig, seed: int = 0, table_std: float = 0.02) -> Dict[str, Tensor]:
    """Deterministic random weights with the reference's parameter names and shapes (ltx.py:291-336,
    transformer.py:70-133).  Tables are non-zero (std `table_std`) so AdaLN shifts/scales/gates are exercised."""
    g = torch.Generator().manual_seed(seed)
    t: Dict[str, Tensor] = {}
    D, Da = cfg.inner_dim, cfg.audio_inner_dim
    vid, aud = cfg.model_type.is_video_enabled(), cfg.model_type.is_audio_enabled()
    if vid:
        _lin(t, "patchify_proj", D, cfg.in_channels, g)
        _adaln(t, "adaln_single", D, 6, g)
        _lin(t, "caption_projection.linear1", D, cfg.caption_channels, g)
        _lin(t, "caption_projection.linear2", D, D, g)
        t["scale_shift_table"] = table_std * torch.randn(2, D, generator=g)
        _lin(t, "proj_out", cfg.out_channels, D, g)
    if aud:
        _lin(t, "audio_patchify_proj", Da, cfg.audio_in_channels, g)
        _adaln(t, "audio_adaln_single", Da, 6, g)
        _lin(t, "audio_caption_projection.linear1", Da, cfg.audio_caption_channels, g)
        _lin(t, "audio_caption_projection.linear2", Da, Da, g)
        t["audio_scale_shift_table"] = table_std * torch.randn(2, Da, generator=g)
        _lin(t, "audio_proj_out", cfg.audio_out_channels, Da, g)
    if vid and aud:
        _adaln(t, "av_ca_video_scale_shift_adaln_single", D, 4, g)
        _adaln(t, "av_ca_audio_scale_shift_adaln_single", Da, 4, g)
        _adaln(t, "av_ca_a2v_gate_adaln_single", D, 1, g)
        _adaln(t, "av_ca_v2a_gate_adaln_single", Da, 1, g)
    for i in range(cfg.num_layers):
        b = f"transformer_blocks.{i}"
        if vid:
            _attn(t, b + ".attn1", D, D, D, g)
            _attn(t, b + ".attn2", D, cfg.cross_attention_dim, D, g)
            _lin(t, b + ".ff.proj_in", 4 * D, D, g)
            _lin(t, b + ".ff.proj_out", D, 4 * D, g)
            t[b + ".scale_shift_table"] = table_std * torch.randn(6, D, generator=g)
        if aud:
            _attn(t, b + ".audio_attn1", Da, Da, Da, g)
            _attn(t, b + ".audio_attn2", Da, cfg.audio_cross_attention_dim, Da, g)
            _lin(t, b + ".audio_ff.proj_in", 4 * Da, Da, g)
            _lin(t, b + ".audio_ff.proj_out", Da, 4 * Da, g)
            t[b + ".audio_scale_shift_table"] = table_std * torch.randn(6, Da, generator=g)
        if vid and aud:
            _attn(t, b + ".audio_to_video_attn", D, Da, Da, g)
            _attn(t, b + ".video_to_audio_attn", Da, D, Da, g)
            t[b + ".scale_shift_table_a2v_ca_audio"] = table_std * torch.randn(5, Da, generator=g)
            t[b + ".scale_shift_table_a2v_ca_video"] = table_std * torch.randn(5, D, generator=g)
    return t


def small_config(model_type=LTXModelType.VideoOnly, num_layers=2, heads=4, audio_heads=4) -> OracleConfig:
    """A reduced-width config with the production structure (head dims 128 / 64 kept) for fast parity tests."""
    return OracleConfig(model_type=model_type, num_attention_heads=heads, attention_head_dim=128, num_layers=num_layers,
                        cross_attention_dim=heads * 128, caption_channels=256, audio_num_attention_heads=audio_heads,
                        audio_attention_head_dim=64, audio_cross_attention_dim=audio_heads * 64, audio_caption_channels=256)


# --------------------------------------------------------------------------------------------------
# LoRA merge into non-quantised base weights (the path generate.py:2997-3007 takes when the
# checkpoint is not quantised).  Restates mlx_video/lora.py:18-129; pinned by tests/golden/lora.npz, which
# oracle/make_golden_lora.py produced by running that file itself.
# --------------------------------------------------------------------------------------------------
def lora_sanitize_prefix(prefix: str) -> str:
    """lora.py:18-33: strip the PyTorch prefixes, apply the renames of LTXModel.sanitize."""
    for p in ("model.diffusion_model.", "diffusion_model."):
        if prefix.startswith(p):
            prefix = prefix[len(p):]
    for old, new in ((".to_out.0.", ".to_out."), (".ff.net.0.proj.", ".ff.proj_in."), (".ff.net.2.", ".ff.proj_out."),
                     (".audio_ff.net.0.proj.", ".audio_ff.proj_in."), (".audio_ff.net.2.", ".audio_ff.proj_out."),
                     (".linear_1.", ".linear1."), (".linear_2.", ".linear2.")):
        prefix = prefix.replace(old, new)
    return prefix


def lora_candidate_keys(base_raw: str, base_sanitized: str) -> List[str]:
    """lora.py:72-90: the weight-dict keys a LoRA pair may refer to, in the reference's order of preference."""
    cand = [base_sanitized, base_raw]
    if base_raw.startswith("diffusion_model."):
        cand.append(f"model.{base_raw}")
    if base_sanitized and not base_sanitized.startswith("model."):
        cand += [f"diffusion_model.{base_sanitized}", f"model.diffusion_model.{base_sanitized}"]
    seen, out = set(), []
    for k in cand:
        if k not in seen:
            seen.add(k)
            out.append(k)
    return out


def lora_pairs(lora_sd: Dict[str, Tensor]):
    """lora.py:56-69: (base key raw, base key sanitised, A (r, in), B (out, r)) in the file's key order."""
    for key in lora_sd:
        if not key.endswith(".lora_A.weight"):
            continue
        key_b = key[: -len(".lora_A.weight")] + ".lora_B.weight"
        if key_b not in lora_sd:
            continue
        base = key.replace(".lora_A.weight", ".weight")
        yield base, lora_sanitize_prefix(base), lora_sd[key], lora_sd[key_b]


def apply_lora_to_weights(weights: Dict[str, Tensor], loras) -> Dict[str, Tensor]:
    """lora.py:93-129.  ``loras``: iterable of (state dict, strength), applied in order:
    delta = (B @ A in fp32) * strength, cast to the weight's dtype, added in the weight's dtype."""
    updated = dict(weights)
    for lora_sd, strength in loras:
        for base_raw, base_san, A, B in lora_pairs(lora_sd):
            key = next((k for k in lora_candidate_keys(base_raw, base_san) if k in updated), None)
            if key is None:
                continue
            w = updated[key]
            delta = (B.float() @ A.float()) * strength
            updated[key] = w + delta.to(w.dtype)
    return updated


# --------------------------------------------------------------------------------------------------
# MLX affine group quantisation (the quantised half).  The reference never does this
# arithmetic itself: ``LTXModel.from_pretrained`` (ltx.py:641-725) swaps the linears named by the checkpoint's
# ``.scales`` tensors for ``nn.QuantizedLinear`` through ``nn.quantize`` and the forward is
# ``mx.quantized_matmul`` — third-party mlx==0.30.1 (uv.lock:760-761), absent here.  Restated from MLX's published
# definition (``mx.quantize`` / ``mx.dequantize`` docs, mode="affine"):
#     w[o, g*G + j] ~= scales[o, g] * q[o, g*G + j] + biases[o, g],      q in [0, 2^bits - 1]
#   packing: 32/bits consecutive q of a row share one uint32, element j of the word in bits [bits*j, bits*(j+1))
#   (lowest bits first) -> weight (out, in*bits/32) uint32, scales / biases (out, in/G) in the weight's dtype.
# ``affine_dequantize`` is the load path and exact by definition.  ``affine_quantize`` only manufactures test
# checkpoints (converting is offline tooling, out of scope); it follows the published fallback algorithm
# (per group: range/(2^bits-1), sign chosen so the larger-magnitude edge is the bias, scale re-fitted so that
# edge/scale is an integer) — parity of the LOAD path does not depend on it.
# --------------------------------------------------------------------------------------------------
def affine_quantize(w: Tensor, group_size: int = 64, bits: int = 4) -> Tuple[np.ndarray, Tensor, Tensor]:
    """-> (packed uint32 ndarray (out, in*bits/32), scales, biases (out, in/group) in w.dtype)."""
    assert w.dim() == 2 and w.shape[1] % group_size == 0 and 32 % bits == 0
    n_bins = float((1 << bits) - 1)
    g = w.reshape(w.shape[0], -1, group_size)
    w_max = g.amax(-1, keepdim=True).float()
    w_min = g.amin(-1, keepdim=True).float()
    mask = w_min.abs() > w_max.abs()
    scales = torch.clamp((w_max - w_min) / n_bins, min=1e-7)
    scales = torch.where(mask, scales, -scales)
    edge = torch.where(mask, w_min, w_max)
    q0 = torch.round(edge / scales)
    scales = torch.where(q0 != 0, edge / torch.where(q0 != 0, q0, torch.ones_like(q0)), scales)
    biases = torch.where(q0 == 0, torch.zeros_like(edge), edge)
    scales, biases = scales.to(w.dtype), biases.to(w.dtype)  # stored in the weight's dtype ...
    q = torch.clamp(torch.round((g.float() - biases.float()) / scales.float()), 0, n_bins)  # ... and used as stored
    per = 32 // bits
    qi = q.reshape(w.shape[0], -1, per).to(torch.int64).numpy().astype(np.uint64)
    shifts = (np.arange(per, dtype=np.uint64) * np.uint64(bits))
    packed = (qi << shifts).sum(-1).astype(np.uint32)
    return packed, scales.squeeze(-1), biases.squeeze(-1)


def affine_unpack(packed: np.ndarray, bits: int) -> np.ndarray:
    """uint32 (out, in*bits/32) -> integer levels (out, in) as int64."""
    per = 32 // bits
    shifts = (np.arange(per, dtype=np.uint32) * np.uint32(bits))
    q = (packed.astype(np.uint32)[..., None] >> shifts) & np.uint32((1 << bits) - 1)
    return q.reshape(packed.shape[0], -1).astype(np.int64)


def affine_dequantize(packed: np.ndarray, scales: Tensor, biases: Tensor, group_size: int = 64, bits: int = 4) -> Tensor:
    """fp32 (out, in): scales * q + biases with scales / biases up-cast to fp32 (one fused multiply-add has no
    intermediate rounding: q <= 255 and an 8-bit-mantissa scale multiply exactly in fp32)."""
    q = torch.from_numpy(affine_unpack(np.asarray(packed), bits)).float()
    out_f, in_f = q.shape
    assert scales.shape == (out_f, in_f // group_size) == biases.shape, (scales.shape, q.shape, group_size)
    s = scales.float().repeat_interleave(group_size, dim=1)
    b = biases.float().repeat_interleave(group_size, dim=1)
    return q * s + b


def quant_params_from_shapes(packed_cols: int, n_groups: int, in_features: int) -> Tuple[int, int]:
    """(group_size, bits) implied by the tensor shapes of one quantised linear."""
    return in_features // n_groups, packed_cols * 32 // in_features


def dequantize_state_dict(weights: Dict[str, object], in_features: Dict[str, int]) -> Dict[str, Tensor]:
    """A state dict holding MLX-quantised linears (``X.weight`` uint32 + ``X.scales`` + ``X.biases``) -> plain fp32
    weights for ``OracleLTXModel``; ``in_features[name]`` is the linear's input width (from the model's shapes)."""
    out: Dict[str, Tensor] = {}
    for k, v in weights.items():
        if k.endswith(".scales") or k.endswith(".biases"):
            continue
        base = k[: -len(".weight")] if k.endswith(".weight") else None
        if base is not None and f"{base}.scales" in weights:
            packed = np.asarray(v)
            s, b = weights[f"{base}.scales"], weights[f"{base}.biases"]
            gs, bits = quant_params_from_shapes(packed.shape[1], s.shape[1], in_features[k])
            out[k] = affine_dequantize(packed, s, b, gs, bits)
        else:
            out[k] = v if isinstance(v, Tensor) else torch.from_numpy(np.asarray(v))
    return out


# --------------------------------------------------------------------------------------------------
# Denoise loops — the direct callers of the forward.  Restate the eager
# (compile_step=False) branches of mlx_video/generate.py: ``denoise_distilled`` :564-881, ``denoise_audio_only``
# :888-1058, ``denoise_dev`` :1060-1327, ``denoise_dev_av`` :1330-1703.  Pinned by tests/golden/sampler.npz, which
# oracle/make_golden_sampler.py produced by running those four functions themselves over the shim.
# Latents keep the reference layouts: video (B, C, F, H, W), audio (B, 8, Ta, 16).
# --------------------------------------------------------------------------------------------------
@dataclass
class LatentState:  # conditioning/latent.py: latent, clean conditioning latent, per-frame mask (B,1,F,1,1), 1 = denoise
    latent: Tensor
    clean_latent: Tensor
    denoise_mask: Tensor


def _video_rope(model: "OracleLTXModel", positions: Tensor):
    c = model.config
    return precompute_freqs_cis(positions, c.inner_dim, c.positional_embedding_theta, list(c.positional_embedding_max_pos),
                                c.use_middle_indices_grid, c.num_attention_heads, c.rope_type, c.double_precision_rope)


def _audio_rope(model: "OracleLTXModel", positions: Tensor):
    c = model.config
    return precompute_freqs_cis(positions, c.audio_inner_dim, c.positional_embedding_theta,
                                list(c.audio_positional_embedding_max_pos), c.use_middle_indices_grid,
                                c.audio_num_attention_heads, c.rope_type, c.double_precision_rope)


def _video_tokens(latents: Tensor) -> Tensor:  # generate.py:792
    b, c = latents.shape[:2]
    return latents.reshape(b, c, -1).transpose(1, 2)


def _audio_tokens(latents: Tensor) -> Tensor:  # generate.py:806-807
    ab, ac, at, af = latents.shape
    return latents.permute(0, 2, 1, 3).reshape(ab, at, ac * af)


def _audio_from_tokens(v: Tensor, shape) -> Tensor:  # generate.py:826-827
    ab, ac, at, af = shape
    return v.reshape(ab, at, ac, af).permute(0, 2, 1, 3)


def _timestep_mask(state: Optional[LatentState], b: int, f: int, h: int, w: int, dtype) -> Tensor:  # generate.py:597-606
    if state is None:
        return torch.ones(b, f * h * w, dtype=dtype)
    return state.denoise_mask.reshape(b, 1, f, 1, 1).expand(b, 1, f, h, w).reshape(b, f * h * w).to(dtype)


def _step(latents: Tensor, denoised: Tensor, sigma: float, sigma_next: float) -> Tensor:  # generate.py:832-846
    return euler_step(latents, denoised, sigma, sigma_next) if sigma_next > 0 else denoised


def denoise_distilled(latents, positions, text_embeddings, model: "OracleLTXModel", sigmas, state: Optional[LatentState] = None,
                      audio_latents=None, audio_positions=None, audio_embeddings=None):
    """generate.py:564-881 (no CFG; optional joint audio)."""
    dtype = latents.dtype
    enable_audio = audio_latents is not None
    if state is not None:
        latents = state.latent
    sig = [float(s) for s in sigmas]
    sig_t = torch.tensor(sig, dtype=dtype)
    b, c, f, h, w = latents.shape
    tmask = _timestep_mask(state, b, f, h, w, dtype)
    rope_v = _video_rope(model, positions)
    if enable_audio:
        if audio_positions is None or audio_embeddings is None:
            raise ValueError("audio_positions/audio_embeddings must be provided when audio_latents is enabled")
        amask = torch.ones(audio_latents.shape[0], audio_latents.shape[2], dtype=dtype)
        rope_a = _audio_rope(model, audio_positions)
    for i in range(len(sig) - 1):
        vm = Modality(_video_tokens(latents), sig_t[i] * tmask, positions, text_embeddings, True, None, rope_v)
        am = Modality(_audio_tokens(audio_latents), sig_t[i] * amask, audio_positions, audio_embeddings, True, None,
                      rope_a) if enable_audio else None
        v, va = model(vm, am)
        den = to_denoised(latents, v.transpose(1, 2).reshape(b, c, f, h, w), sig_t[i])
        if state is not None:
            den = apply_denoise_mask(den, state.clean_latent, state.denoise_mask)
        latents = _step(latents, den, sig[i], sig[i + 1])
        if enable_audio and va is not None:
            aden = to_denoised(audio_latents, _audio_from_tokens(va, audio_latents.shape), sig_t[i])
            audio_latents = _step(audio_latents, aden, sig[i], sig[i + 1])
    return latents, (audio_latents if enable_audio else None)


def denoise_audio_only(audio_latents, audio_positions, audio_embeddings, model: "OracleLTXModel", sigmas):
    """generate.py:888-1058 (AudioOnly transformer, no CFG)."""
    dtype = audio_latents.dtype
    sig = [float(s) for s in sigmas]
    sig_t = torch.tensor(sig, dtype=dtype)
    amask = torch.ones(audio_latents.shape[0], audio_latents.shape[2], dtype=dtype)
    rope_a = _audio_rope(model, audio_positions)
    for i in range(len(sig) - 1):
        am = Modality(_audio_tokens(audio_latents), sig_t[i] * amask, audio_positions, audio_embeddings, True, None, rope_a)
        _, va = model(None, am)
        aden = to_denoised(audio_latents, _audio_from_tokens(va, audio_latents.shape), sig_t[i])
        audio_latents = _step(audio_latents, aden, sig[i], sig[i + 1])
    return audio_latents


def denoise_dev(latents, positions, text_embeddings_pos, text_embeddings_neg, model: "OracleLTXModel", sigmas,
                cfg_scale: float = 4.0, state: Optional[LatentState] = None, cfg_batch: bool = False):
    """generate.py:1060-1327 (CFG; ``cfg_batch`` = cond and uncond as one B=2 forward, :1239-1255)."""
    dtype = latents.dtype
    if state is not None:
        latents = state.latent
    sig = [float(s) for s in (sigmas.tolist() if hasattr(sigmas, "tolist") else sigmas)]
    sig_t = torch.tensor(sig, dtype=torch.float32).to(dtype)
    use_cfg = cfg_scale != 1.0
    cfg_batch = cfg_batch and use_cfg
    b, c, f, h, w = latents.shape
    tmask = _timestep_mask(state, b, f, h, w, dtype)
    rope = _video_rope(model, positions)
    for i in range(len(sig) - 1):
        flat, ts = _video_tokens(latents), sig_t[i] * tmask
        if cfg_batch:
            vv, _ = model(Modality(torch.cat([flat, flat], 0), torch.cat([ts, ts], 0), positions.expand(2 * b, *positions.shape[1:]),
                                   torch.cat([text_embeddings_pos, text_embeddings_neg], 0), True, None,
                                   tuple(t.expand(2 * b, *t.shape[1:]) for t in rope)), None)
            v = cfg_combine(vv[:b], vv[b:], cfg_scale)
        else:
            v, _ = model(Modality(flat, ts, positions, text_embeddings_pos, True, None, rope), None)
            if use_cfg:
                vn, _ = model(Modality(flat, ts, positions, text_embeddings_neg, True, None, rope), None)
                v = cfg_combine(v, vn, cfg_scale)
        den = to_denoised(latents, v.transpose(1, 2).reshape(b, c, f, h, w), sig_t[i])
        if state is not None:
            den = apply_denoise_mask(den, state.clean_latent, state.denoise_mask)
        latents = _step(latents, den, sig[i], sig[i + 1])
    return latents


def denoise_dev_av(video_latents, audio_latents, video_positions, audio_positions, video_embeddings_pos, video_embeddings_neg,
                   audio_embeddings_pos, audio_embeddings_neg, model: "OracleLTXModel", sigmas, cfg_scale: float = 4.0,
                   video_state: Optional[LatentState] = None, cfg_batch: bool = False):
    """generate.py:1330-1703 (CFG on both modalities of the joint audio+video model)."""
    dtype = video_latents.dtype
    if video_state is not None:
        video_latents = video_state.latent
    sig = [float(s) for s in (sigmas.tolist() if hasattr(sigmas, "tolist") else sigmas)]
    sig_t = torch.tensor(sig, dtype=torch.float32).to(dtype)
    use_cfg = cfg_scale != 1.0
    cfg_batch = cfg_batch and use_cfg
    b, c, f, h, w = video_latents.shape
    ab = audio_latents.shape[0]
    tmask = _timestep_mask(video_state, b, f, h, w, dtype)
    amask = torch.ones(ab, audio_latents.shape[2], dtype=dtype)
    rope_v, rope_a = _video_rope(model, video_positions), _audio_rope(model, audio_positions)

    def twice(t: Tensor) -> Tensor:
        return torch.cat([t, t], 0)

    for i in range(len(sig) - 1):
        vflat, aflat = _video_tokens(video_latents), _audio_tokens(audio_latents)
        vts, ats = sig_t[i] * tmask, sig_t[i] * amask
        if cfg_batch:
            vv, av = model(Modality(twice(vflat), twice(vts), video_positions.expand(2 * b, *video_positions.shape[1:]),
                                    torch.cat([video_embeddings_pos, video_embeddings_neg], 0), True, None,
                                    tuple(t.expand(2 * b, *t.shape[1:]) for t in rope_v)),
                           Modality(twice(aflat), twice(ats), audio_positions.expand(2 * ab, *audio_positions.shape[1:]),
                                    torch.cat([audio_embeddings_pos, audio_embeddings_neg], 0), True, None,
                                    tuple(t.expand(2 * ab, *t.shape[1:]) for t in rope_a)))
            v, va = cfg_combine(vv[:b], vv[b:], cfg_scale), cfg_combine(av[:ab], av[ab:], cfg_scale)
        else:
            v, va = model(Modality(vflat, vts, video_positions, video_embeddings_pos, True, None, rope_v),
                          Modality(aflat, ats, audio_positions, audio_embeddings_pos, True, None, rope_a))
            if use_cfg:
                vn, van = model(Modality(vflat, vts, video_positions, video_embeddings_neg, True, None, rope_v),
                                Modality(aflat, ats, audio_positions, audio_embeddings_neg, True, None, rope_a))
                v, va = cfg_combine(v, vn, cfg_scale), cfg_combine(va, van, cfg_scale)
        vden = to_denoised(video_latents, v.transpose(1, 2).reshape(b, c, f, h, w), sig_t[i])
        aden = to_denoised(audio_latents, _audio_from_tokens(va, audio_latents.shape), sig_t[i])
        if video_state is not None:
            vden = apply_denoise_mask(vden, video_state.clean_latent, video_state.denoise_mask)
        video_latents = _step(video_latents, vden, sig[i], sig[i + 1])
        audio_latents = _step(audio_latents, aden, sig[i], sig[i + 1])
    return video_latents, audio_latents
