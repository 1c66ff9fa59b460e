"""Sampler side of the denoise step: position grids, sigma schedules and the denoise loops that own
the per-step call into ``LTXModel``.

Mirrors the reference's ``mlx_video/generate.py``: ``create_position_grid`` :470-525,
``create_audio_position_grid`` :528-551, ``compute_audio_frames`` :554-557, ``ltx2_scheduler`` :410-467,
constants :339-353, ``denoise_distilled`` :564-881, ``denoise_audio_only`` :888-1058, ``denoise_dev`` :1060-1327,
``denoise_dev_av`` :1330-1703 (same argument names and meaning; pinned by tests/golden/sampler.npz, which the
reference's own four loops produced).  Grids and schedules are host numpy like the reference (integer / index work: bit-exact).
The per-step latent update — CFG combine, x0 = x - sigma v, conditioning-mask blend, fp32 Euler
(utils.py:404-440; generate.py:1255,1283-1301) — is one fused kernel (``ltxb_euler_step``) on fp32
latents kept token-major (B, T, C) for the whole loop.

Deliberate deviation ("reference bf16 quirk"): the loops keep fp32 latents and fp32 sigmas for the whole
loop and cast to the input dtype once at the end.  The reference builds ``sigmas_mx`` in the latents' dtype and casts the
latents back every step (generate.py:592,653,846), so a bf16 run there feeds bf16-rounded timesteps (0.99375 becomes
0.9921875) and rounds the latents 8 / 40 times; the golden fixture (tests/golden/sampler.npz) pins the fp32 behaviour, where
both agree bit for bit.  With bf16 inputs the results here are the fp32-loop results rounded once (closer to the fp32
reference than the reference's own bf16 run).
"""
from __future__ import annotations

import math
import os
from contextlib import contextmanager
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops
from ._lib import LtxbError
from .model import LTXModel
from .rope import precompute_freqs_cis
from .transformer import Modality

Tensor = torch.Tensor

STAGE_1_SIGMAS = [1.0, 0.99375, 0.9875, 0.98125, 0.975, 0.909375, 0.725, 0.421875, 0.0]
STAGE_2_SIGMAS = [0.909375, 0.725, 0.421875, 0.0]
BASE_SHIFT_ANCHOR = 1024
MAX_SHIFT_ANCHOR = 4096
AUDIO_SAMPLE_RATE = 24000
AUDIO_LATENT_SAMPLE_RATE = 16000
AUDIO_HOP_LENGTH = 160
AUDIO_LATENT_DOWNSAMPLE_FACTOR = 4
AUDIO_LATENT_CHANNELS = 8
AUDIO_MEL_BINS = 16
AUDIO_LATENTS_PER_SECOND = AUDIO_LATENT_SAMPLE_RATE / AUDIO_HOP_LENGTH / AUDIO_LATENT_DOWNSAMPLE_FACTOR


def create_position_grid(batch_size: int, num_frames: int, height: int, width: int, temporal_scale: int = 8,
                         spatial_scale: int = 32, fps: float = 24.0, causal_fix: bool = True) -> np.ndarray:
    """(B, 3, T, 2) fp32 [start, end) bounds in (seconds, px, px); token t = f*H*W + h*W + w.
    Integer patch index -> pixel via x8 / x32 / x32, causal first-frame fix max(0, p + 1 - 8) on the time
    axis, then /fps — the same fp32 rounding sequence as generate.py:470-525."""
    T = num_frames * height * width
    idx = np.arange(T, dtype=np.int64)
    per_axis = np.stack([idx // (height * width), (idx // width) % height, idx % width])  # (3, T) integer
    bounds = np.stack([per_axis, per_axis + 1], axis=-1)  # (3, T, 2)
    scale = np.array([temporal_scale, spatial_scale, spatial_scale], dtype=np.int64).reshape(3, 1, 1)
    px = (bounds * scale).astype(np.float32)
    if causal_fix:
        px[0] = np.clip(px[0] + 1 - temporal_scale, a_min=0, a_max=None)
    px[0] = px[0] / fps
    return np.array(np.broadcast_to(px[None], (batch_size, 3, T, 2)), dtype=np.float32, order="C", copy=True)


def create_audio_position_grid(batch_size: int, audio_frames: int, sample_rate: int = AUDIO_LATENT_SAMPLE_RATE,
                               hop_length: int = AUDIO_HOP_LENGTH,
                               downsample_factor: int = AUDIO_LATENT_DOWNSAMPLE_FACTOR,
                               is_causal: bool = True) -> np.ndarray:
    """(B, 1, Ta, 2) fp32 seconds: latent frame -> mel frame (x4, causal fix) -> seconds (generate.py:528-551)."""
    def seconds(first: int) -> np.ndarray:
        mel = np.arange(first, first + audio_frames, dtype=np.float32) * downsample_factor
        if is_causal:
            mel = np.clip(mel + 1 - downsample_factor, 0, None)
        return mel * hop_length / sample_rate

    pos = np.stack([seconds(0), seconds(1)], axis=-1)[None, None]
    return np.array(np.broadcast_to(pos, (batch_size, 1, audio_frames, 2)), dtype=np.float32, order="C", copy=True)


def compute_audio_frames(num_video_frames: int, fps: float) -> int:
    """generate.py:554-557"""
    return round(num_video_frames / fps * AUDIO_LATENTS_PER_SECOND)


def ltx2_scheduler(steps: int, num_tokens: Optional[int] = None, max_shift: float = 2.05, base_shift: float = 0.95,
                   stretch: bool = True, terminal: float = 0.1) -> np.ndarray:
    """Token-count-dependent shifted sigma schedule, stretched to end at ``terminal`` (generate.py:410-467).
    fp64 numpy arithmetic, fp32 result, like the reference."""
    tokens = MAX_SHIFT_ANCHOR if num_tokens is None else min(num_tokens, MAX_SHIFT_ANCHOR)
    slope = (max_shift - base_shift) / (MAX_SHIFT_ANCHOR - BASE_SHIFT_ANCHOR)
    shift = tokens * slope + (base_shift - slope * BASE_SHIFT_ANCHOR)
    lin = np.linspace(1.0, 0.0, steps + 1)
    sig = np.zeros_like(lin)
    live = lin != 0
    if np.any(live):
        sig[live] = math.exp(shift) / (math.exp(shift) + (1 / lin[live] - 1) ** 1)
    if stretch:
        live = sig != 0
        gap = 1.0 - sig[live]
        factor = gap[-1] / (1.0 - terminal)
        if np.isfinite(factor) and factor != 0:
            sig[live] = 1.0 - gap / factor
    return sig.astype(np.float32)


@dataclass
class LatentState:
    """conditioning/latent.py: latent being denoised, clean conditioning latent, per-frame denoise mask
    (B, 1, F, 1, 1): 1 = denoise, 0 = keep clean."""

    latent: Tensor
    clean_latent: Tensor
    denoise_mask: Tensor


def _to_tokens(latents: Tensor) -> Tensor:
    """(B, C, F, H, W) -> fp32 (B, T, C) token-major (generate.py:792: reshape + transpose). Layout plumbing."""
    b, c = latents.shape[:2]
    return latents.reshape(b, c, -1).transpose(1, 2).to(torch.float32).contiguous()


def _from_tokens(tokens: Tensor, shape, dtype) -> Tensor:
    b, c = shape[:2]
    return tokens.transpose(1, 2).reshape(shape).to(dtype).contiguous()


def _token_mask(state: Optional[LatentState], b: int, f: int, h: int, w: int, device) -> Optional[Tensor]:
    if state is None:
        return None
    m = state.denoise_mask.reshape(b, 1, f, 1, 1).to(device=device, dtype=torch.float32)
    return m.expand(b, 1, f, h, w).reshape(b, f * h * w).contiguous()


def _video_rope(transformer: LTXModel, positions: Tensor):
    return precompute_freqs_cis(positions, dim=transformer.inner_dim, theta=transformer.positional_embedding_theta,
                                max_pos=transformer.positional_embedding_max_pos,
                                use_middle_indices_grid=transformer.use_middle_indices_grid,
                                num_attention_heads=transformer.num_attention_heads, rope_type=transformer.rope_type,
                                double_precision=transformer.config.double_precision_rope)


def _audio_rope(transformer: LTXModel, positions: Tensor):
    return precompute_freqs_cis(positions, dim=transformer.audio_inner_dim, theta=transformer.positional_embedding_theta,
                                max_pos=transformer.audio_positional_embedding_max_pos,
                                use_middle_indices_grid=transformer.use_middle_indices_grid,
                                num_attention_heads=transformer.audio_num_attention_heads,
                                rope_type=transformer.rope_type, double_precision=transformer.config.double_precision_rope)


def _dev(t, device) -> Tensor:
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(t)
    return t.to(device)


@contextmanager
def _loop_context_cache(transformer: LTXModel):
    """The prompt embeddings are arguments of the denoise LOOP (generate.py:564-575, 1060-1075): they cannot change
    between its steps, so the caption projection and every block's text K/V are computed on the first step and reused
    by the others (bit-exact, tests/test_gpu_parity.py).  The reference recomputes them every step;
    LTXB_LOOP_CONTEXT_CACHE=0 does the same here."""
    prev = transformer.cache_context
    if os.environ.get("LTXB_LOOP_CONTEXT_CACHE", "1") != "0":
        transformer.cache_context = True
        transformer.invalidate_context()
    try:
        yield
    finally:
        transformer.cache_context = prev


def _loop_steps(transformer: LTXModel, n: int):
    """range(n) with the loop-level context cache switched on for as long as the loop runs."""
    with _loop_context_cache(transformer):
        yield from range(n)


def _advance(x: Tensor, v_pos: Tensor, sigma: float, sigma_next: float, v_neg: Optional[Tensor] = None,
             cfg_scale: float = 1.0, mask: Optional[Tensor] = None, clean: Optional[Tensor] = None) -> None:
    """One fused latent update on fp32 tokens (B, T, C), in place.  sigma_next == 0 gives x = x0 exactly as the
    reference's ``latents = denoised`` branch (generate.py:1302)."""
    B, T, C = x.shape
    ops.euler_step(x.view(B * T, C), v_pos.reshape(B * T, C), float(sigma), float(sigma_next),
                   v_neg=None if v_neg is None else v_neg.reshape(B * T, C), cfg_scale=float(cfg_scale),
                   mask=None if mask is None else mask.reshape(B * T), clean=None if clean is None else clean.view(B * T, C))


def denoise_distilled(latents: Tensor, positions, text_embeddings: Tensor, transformer: LTXModel, sigmas: Sequence[float],
                      verbose: bool = False, state: Optional[LatentState] = None, audio_latents: Optional[Tensor] = None,
                      audio_positions=None, audio_embeddings: Optional[Tensor] = None, eval_interval: int = 1,
                      compile_step: bool = False, compile_shapeless: bool = False, fp32_euler: bool = True,
                      ui_phase: str = "denoise") -> Tuple[Tensor, Optional[Tensor]]:
    """generate.py:564-881 — distilled pipeline, no CFG, one forward per step.  latents (B, C, F, H, W);
    audio_latents (B, 8, Ta, 16).  Work is stream-ordered; nothing synchronises inside the loop
    (``eval_interval`` / ``compile_*`` are accepted for signature compatibility)."""
    dev = transformer.device
    dtype = latents.dtype
    if state is not None:
        latents = state.latent
    b, c, f, h, w = latents.shape
    T = f * h * w
    enable_audio = audio_latents is not None
    sig = [float(s) for s in (sigmas.tolist() if hasattr(sigmas, "tolist") else sigmas)]
    positions = _dev(positions, dev)
    x = _to_tokens(latents.to(dev))
    mask = _token_mask(state, b, f, h, w, dev)
    clean = _to_tokens(state.clean_latent.to(dev)) if state is not None else None
    ts_mask = torch.ones(b, T, dtype=torch.float32, device=dev) if mask is None else mask
    rope_v = _video_rope(transformer, positions)
    xa = rope_a = None
    if enable_audio:
        if audio_positions is None or audio_embeddings is None:
            raise ValueError("audio_positions/audio_embeddings must be provided when audio_latents is enabled")
        ab, ac, at, af = audio_latents.shape
        audio_positions = _dev(audio_positions, dev)
        xa = audio_latents.to(dev).permute(0, 2, 1, 3).reshape(ab, at, ac * af).to(torch.float32).contiguous()
        a_ones = torch.ones(ab, at, dtype=torch.float32, device=dev)
        rope_a = _audio_rope(transformer, audio_positions)
    text_embeddings = text_embeddings.to(dev)
    for i in _loop_steps(transformer, len(sig) - 1):
        sigma, sigma_next = sig[i], sig[i + 1]
        vm = Modality(latent=x, timesteps=ts_mask * sigma, positions=positions, context=text_embeddings,
                      context_mask=None, enabled=True, positional_embeddings=rope_v)
        am = None
        if enable_audio:
            am = Modality(latent=xa, timesteps=a_ones * sigma, positions=audio_positions, context=audio_embeddings.to(dev),
                          context_mask=None, enabled=True, positional_embeddings=rope_a)
        v, va = transformer(video=vm, audio=am)
        _advance(x, v, sigma, sigma_next, mask=mask, clean=clean)
        if enable_audio and va is not None:
            _advance(xa, va, sigma, sigma_next)
    out = _from_tokens(x, (b, c, f, h, w), dtype)
    out_a = None
    if enable_audio:
        out_a = xa.reshape(ab, at, ac, af).permute(0, 2, 1, 3).to(audio_latents.dtype).contiguous()
    return out, out_a


def denoise_dev(latents: Tensor, positions, text_embeddings_pos: Tensor, text_embeddings_neg: Tensor,
                transformer: LTXModel, sigmas, cfg_scale: float = 4.0, verbose: bool = False,
                state: Optional[LatentState] = None, eval_interval: int = 1, compile_step: bool = False,
                compile_shapeless: bool = False, cfg_batch: bool = False, ui_phase: str = "denoise",
                cfg_parallel=None) -> Tensor:
    """generate.py:1060-1327 — dev pipeline with classifier-free guidance.  ``cfg_batch`` runs cond and
    uncond as one B=2 forward (generate.py:1239-1255); ``cfg_parallel`` (parallel.CFGParallel) runs them on
    two rank groups and exchanges the velocities over NCCL."""
    dev = transformer.device
    dtype = latents.dtype
    if state is not None:
        latents = state.latent
    sig = [float(s) for s in (sigmas.tolist() if hasattr(sigmas, "tolist") else sigmas)]
    use_cfg = cfg_scale != 1.0
    cfg_batch = cfg_batch and use_cfg and cfg_parallel is None
    b, c, f, h, w = latents.shape
    T = f * h * w
    positions = _dev(positions, dev)
    x = _to_tokens(latents.to(dev))
    mask = _token_mask(state, b, f, h, w, dev)
    clean = _to_tokens(state.clean_latent.to(dev)) if state is not None else None
    ts_mask = torch.ones(b, T, dtype=torch.float32, device=dev) if mask is None else mask
    rope = _video_rope(transformer, positions)
    pos_ctx, neg_ctx = text_embeddings_pos.to(dev), text_embeddings_neg.to(dev)
    if cfg_batch:
        ctx_cat = torch.cat([pos_ctx, neg_ctx], dim=0)
        positions_cfg = positions.expand(2 * b, *positions.shape[1:]) if positions.shape[0] == b and b == 1 else torch.cat([positions, positions], 0)
        rope_cfg = rope if rope[0].shape[0] == 1 else (torch.cat([rope[0]] * 2, 0), torch.cat([rope[1]] * 2, 0))
    for i in _loop_steps(transformer, len(sig) - 1):
        sigma, sigma_next = sig[i], sig[i + 1]
        ts = ts_mask * sigma
        v_neg = None
        if cfg_parallel is not None and use_cfg:
            ctx = pos_ctx if cfg_parallel.is_cond else neg_ctx
            mine, _ = transformer(video=Modality(x, ts, positions, ctx, True, None, rope), audio=None)
            v_pos, v_neg = cfg_parallel.exchange(mine)
        elif cfg_batch:
            xx = torch.cat([x, x], dim=0)
            vv, _ = transformer(video=Modality(xx, torch.cat([ts, ts], 0), positions_cfg, ctx_cat, True, None, rope_cfg), audio=None)
            v_pos, v_neg = vv[:b], vv[b:]
        else:
            v_pos, _ = transformer(video=Modality(x, ts, positions, pos_ctx, True, None, rope), audio=None)
            if use_cfg:
                v_neg, _ = transformer(video=Modality(x, ts, positions, neg_ctx, True, None, rope), audio=None)
        _advance(x, v_pos.contiguous(), sigma, sigma_next, v_neg=None if v_neg is None else v_neg.contiguous(),
                 cfg_scale=cfg_scale, mask=mask, clean=clean)
    return _from_tokens(x, (b, c, f, h, w), dtype)


def _audio_to_tokens(audio_latents: Tensor) -> Tensor:
    """(B, 8, Ta, 16) -> fp32 (B, Ta, 128) token-major (generate.py:806-807). Layout plumbing."""
    ab, ac, at, af = audio_latents.shape
    return audio_latents.permute(0, 2, 1, 3).reshape(ab, at, ac * af).to(torch.float32).contiguous()


def _audio_from_tokens(tokens: Tensor, shape, dtype) -> Tensor:
    ab, ac, at, af = shape
    return tokens.reshape(ab, at, ac, af).permute(0, 2, 1, 3).to(dtype).contiguous()


def denoise_audio_only(audio_latents: Tensor, audio_positions, audio_embeddings: Tensor, transformer: LTXModel,
                       sigmas: Sequence[float], verbose: bool = False, eval_interval: int = 1, compile_step: bool = False,
                       compile_shapeless: bool = False, fp32_euler: bool = True, ui_phase: str = "audio_denoise") -> Tensor:
    """generate.py:888-1058 — distilled-style loop of an AudioOnly transformer (no CFG).  audio_latents (B, 8, Ta, 16)."""
    dev = transformer.device
    sig = [float(s) for s in (sigmas.tolist() if hasattr(sigmas, "tolist") else sigmas)]
    audio_positions = _dev(audio_positions, dev)
    xa = _audio_to_tokens(audio_latents.to(dev))
    ones = torch.ones(xa.shape[0], xa.shape[1], dtype=torch.float32, device=dev)
    rope_a = _audio_rope(transformer, audio_positions)
    ctx = audio_embeddings.to(dev)
    for i in _loop_steps(transformer, len(sig) - 1):
        am = Modality(latent=xa, timesteps=ones * sig[i], positions=audio_positions, context=ctx, context_mask=None,
                      enabled=True, positional_embeddings=rope_a)
        _, va = transformer(video=None, audio=am)
        _advance(xa, va.contiguous(), sig[i], sig[i + 1])
    return _audio_from_tokens(xa, audio_latents.shape, audio_latents.dtype)


def denoise_dev_av(video_latents: Tensor, audio_latents: Tensor, video_positions, audio_positions,
                   video_embeddings_pos: Tensor, video_embeddings_neg: Tensor, audio_embeddings_pos: Tensor,
                   audio_embeddings_neg: Tensor, transformer: LTXModel, sigmas, cfg_scale: float = 4.0, verbose: bool = False,
                   video_state: Optional[LatentState] = None, eval_interval: int = 1, compile_step: bool = False,
                   compile_shapeless: bool = False, cfg_batch: bool = False, ui_phase: str = "denoise",
                   cfg_parallel=None) -> Tuple[Tensor, Tensor]:
    """generate.py:1330-1703 — dev pipeline of the joint audio+video model: classifier-free guidance on BOTH
    velocities.  ``cfg_batch``: cond and uncond as one B=2 forward (:1572-1598); ``cfg_parallel``
    (parallel.CFGParallel): on two rank groups, the two velocities exchanged over NCCL."""
    dev = transformer.device
    dtype = video_latents.dtype
    if video_state is not None:
        video_latents = video_state.latent
    sig = [float(s) for s in (sigmas.tolist() if hasattr(sigmas, "tolist") else sigmas)]
    use_cfg = cfg_scale != 1.0
    cfg_batch = cfg_batch and use_cfg and cfg_parallel is None
    b, c, f, h, w = video_latents.shape
    T = f * h * w
    video_positions, audio_positions = _dev(video_positions, dev), _dev(audio_positions, dev)
    x = _to_tokens(video_latents.to(dev))
    xa = _audio_to_tokens(audio_latents.to(dev))
    ab, at = xa.shape[:2]
    mask = _token_mask(video_state, b, f, h, w, dev)
    clean = _to_tokens(video_state.clean_latent.to(dev)) if video_state is not None else None
    ts_mask = torch.ones(b, T, dtype=torch.float32, device=dev) if mask is None else mask
    a_ones = torch.ones(ab, at, dtype=torch.float32, device=dev)
    rope_v, rope_a = _video_rope(transformer, video_positions), _audio_rope(transformer, audio_positions)
    vp, vn = video_embeddings_pos.to(dev), video_embeddings_neg.to(dev)
    ap, an = audio_embeddings_pos.to(dev), audio_embeddings_neg.to(dev)

    def twice(t: Tensor) -> Tensor:
        return torch.cat([t, t], dim=0)

    if cfg_batch:
        vctx, actx = torch.cat([vp, vn], 0), torch.cat([ap, an], 0)
        vpos2, apos2 = twice(video_positions), twice(audio_positions)
        rope_v2 = rope_v if rope_v[0].shape[0] == 1 else (twice(rope_v[0]), twice(rope_v[1]))
        rope_a2 = rope_a if rope_a[0].shape[0] == 1 else (twice(rope_a[0]), twice(rope_a[1]))

    def forward(vctx_, actx_):
        return transformer(video=Modality(x, ts, video_positions, vctx_, True, None, rope_v),
                           audio=Modality(xa, ats, audio_positions, actx_, True, None, rope_a))

    for i in _loop_steps(transformer, len(sig) - 1):
        sigma, sigma_next = sig[i], sig[i + 1]
        ts, ats = ts_mask * sigma, a_ones * sigma
        v_neg = a_neg = None
        if cfg_parallel is not None and use_cfg:
            mine_v, mine_a = forward(vp, ap) if cfg_parallel.is_cond else forward(vn, an)
            v_pos, v_neg = cfg_parallel.exchange(mine_v)
            a_pos, a_neg = cfg_parallel.exchange(mine_a)
        elif cfg_batch:
            vv, aa = transformer(video=Modality(twice(x), twice(ts), vpos2, vctx, True, None, rope_v2),
                                 audio=Modality(twice(xa), twice(ats), apos2, actx, True, None, rope_a2))
            v_pos, v_neg, a_pos, a_neg = vv[:b], vv[b:], aa[:ab], aa[ab:]
        else:
            v_pos, a_pos = forward(vp, ap)
            if use_cfg:
                v_neg, a_neg = forward(vn, an)
        _advance(x, v_pos.contiguous(), sigma, sigma_next, v_neg=None if v_neg is None else v_neg.contiguous(),
                 cfg_scale=cfg_scale, mask=mask, clean=clean)
        _advance(xa, a_pos.contiguous(), sigma, sigma_next, v_neg=None if a_neg is None else a_neg.contiguous(),
                 cfg_scale=cfg_scale)
    return _from_tokens(x, (b, c, f, h, w), dtype), _audio_from_tokens(xa, audio_latents.shape, audio_latents.dtype)
