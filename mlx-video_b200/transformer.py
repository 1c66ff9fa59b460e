"""The LTX-2 DiT transformer block on sm_100a kernels.

Host-side mirror of the reference's ``mlx_video/models/ltx/transformer.py`` (``Modality`` :13-22,
``TransformerArgs`` :25-36, ``BasicAVTransformerBlock`` :38-361), ``attention.py:56-142`` and
``feed_forward.py:17-40`` — same class names, constructor arguments, attribute names and call
signatures — with every array op replaced by a call into ``libltxb.so`` (include/ltxb.h).

Layout in HBM (one block, D = inner dim, M = B*T rows):
  residual stream x           f32  [M, D]      updated in place by the GEMM epilogues
  normalised / modulated nx   bf16 [M, D]      A operand of the next GEMM
  qkv                         bf16 [M, 3D]     one GEMM against the row-concatenated to_q|to_k|to_v weight;
                                               q, k, v are column slices (leading dimension 3D)
  attention output o          bf16 [M, D]
  FFN hidden                  bf16 [M, 4D]     GELU-tanh applied in the GEMM epilogue
  modulation rows             f32  [R, 6D]     R = distinct timesteps (or B, or M), addressed per token
                                               through ``timestep_index`` / a row divisor
Weights are bf16 ``(out, in)`` like nn.Linear; biases, norm weights and scale-shift tables are f32.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, replace
from typing import Dict, Iterator, Optional, Tuple

import torch

from . import _lib, ops
from .config import LTXRopeType, TransformerConfig

Tensor = torch.Tensor
BF16, F32 = torch.bfloat16, torch.float32
# A/B switch (scripts/bench_ab.sh): 0 = the attention out-projections add their residual in the GEMM epilogue
# (LTXB_EPI_RESID_GATE_F32) instead of handing it to the norm kernel that follows
DEFER_RESIDUAL = os.environ.get("LTXB_DEFER_RESID", "1") != "0"


@dataclass(frozen=True)
class Modality:
    """Input record of one modality (transformer.py:13-22)."""

    latent: Tensor  # (B, T, C)
    timesteps: Tensor  # (B, T) per-token sigma, or (B, 1)
    positions: Tensor  # (B, 3, T, 2) video / (B, 1, T, 2) audio, fp32 [start, end) bounds
    context: Tensor  # (B, Tc, caption_channels)
    enabled: bool = True
    context_mask: Optional[Tensor] = None
    positional_embeddings: Optional[Tuple[Tensor, Tensor]] = None  # precomputed (cos, sin)


@dataclass(frozen=True)
class TransformerArgs:
    """State threaded through the blocks (transformer.py:25-36).

    ``timesteps`` / ``embedded_timestep`` / the cross tensors hold one row per (batch, token) —
    shape (B, T, k*D) — or per batch — (B, 1, k*D) — exactly as in the reference, or, when the model
    deduplicated the timesteps, (1, R, k*D) with ``timestep_index`` (int32 [B*T]) naming each token's row.
    """

    x: Tensor  # f32 (B, T, D)
    context: Tensor  # bf16 (B, Tc, D)
    context_mask: Optional[Tensor]  # f32 additive bias (B, Tc), or None
    timesteps: Tensor  # f32 (*, *, 6D)
    embedded_timestep: Tensor  # f32 (*, *, D)
    positional_embeddings: Optional[Tuple[Tensor, Tensor]]
    cross_positional_embeddings: Optional[Tuple[Tensor, Tensor]] = None
    cross_scale_shift_timestep: Optional[Tensor] = None
    cross_gate_timestep: Optional[Tensor] = None
    enabled: bool = True
    timestep_index: Optional[Tensor] = None
    context_cache: Optional["ContextCache"] = None  # opt-in reuse of the text K/V across denoise steps
    # text-cross-attention K|V of ALL blocks, k-norm applied: bf16 (B*Tc, L*2*inner), block l at columns
    # [l*2*inner, (l+1)*2*inner) — one GEMM + one norm launch at the head of the forward (model.py) instead of 48 + 48
    text_kv: Optional[Tensor] = None


class ContextCache:
    """Projected caption and per-block text-cross-attention K/V (after k-norm) of ONE context tensor.  The
    context, hence these projections, are constant across all steps of a denoise loop;
    the reference recomputes them every step.  Buffers keep their addresses when the context changes (only
    ``valid`` drops), so captured CUDA graphs that fill or read them stay correct."""

    def __init__(self) -> None:
        self.key = None
        self.stamp = 0  # last use (LTXModel keeps a few of these and re-targets the least recently used)
        self.context: Optional[Tensor] = None  # projected caption, bf16 (B, Tc, D)
        self.kv: Dict[object, Tensor] = {}
        self.valid = False

    def retarget(self, key) -> None:
        if key != self.key:
            self.key, self.valid = key, False

    def context_buffer(self, shape, device) -> Tensor:
        if self.context is None or tuple(self.context.shape) != tuple(shape):
            self.context, self.valid = torch.empty(shape, dtype=BF16, device=device), False
        return self.context

    def stacked(self, shape, device) -> Tensor:
        """K|V of all blocks in one buffer (TransformerArgs.text_kv)."""
        t = self.kv.get("all")
        if t is None or tuple(t.shape) != tuple(shape):
            t = self.kv["all"] = torch.empty(shape, dtype=BF16, device=device)
            self.valid = False
        return t

    def entry(self, block_idx: int, shape, device) -> Tensor:
        t = self.kv.get(block_idx)
        if t is None or tuple(t.shape) != tuple(shape):
            t = self.kv[block_idx] = torch.empty(shape, dtype=BF16, device=device)
            self.valid = False
        return t


class Workspace:
    """Named scratch buffers, allocated once per (tag, shape) and reused by all 48 blocks, so the
    forward is allocation-free after the first call (and its pointers are stable for CUDA graphs)."""

    def __init__(self) -> None:
        self._bufs: Dict[tuple, Tensor] = {}

    def get(self, tag: str, shape, dtype, device) -> Tensor:
        key = (tag, tuple(int(s) for s in shape), dtype, str(device))
        buf = self._bufs.get(key)
        if buf is None:
            buf = torch.empty(key[1], dtype=dtype, device=device)
            self._bufs[key] = buf
        return buf

    def clear(self) -> None:
        self._bufs.clear()

    def nbytes(self) -> int:
        return sum(b.numel() * b.element_size() for b in self._bufs.values())


class Linear:
    """nn.Linear parameters: ``weight`` bf16 (out, in) — possibly a row slice of a fused matrix — and
    ``bias`` f32 (out,)."""

    def __init__(self, weight: Tensor, bias: Tensor) -> None:
        assert weight.dtype == BF16 and bias.dtype == F32 and weight.shape[0] == bias.shape[0]
        self.weight, self.bias = weight, bias

    def __call__(self, x: Tensor, out: Optional[Tensor] = None, mode: int = _lib.EPI_BIAS_BF16) -> Tensor:
        x2 = x.reshape(-1, x.shape[-1])
        if out is None:
            dt = F32 if mode == _lib.EPI_BIAS_F32 else BF16
            out = torch.empty((*x.shape[:-1], self.weight.shape[0]), dtype=dt, device=x.device)
        ops.gemm(x2, self.weight, self.bias, out.view(-1, out.shape[-1]), mode, const_w=True)
        return out


class RMSNormWeight:
    """Holder of an nn.RMSNorm weight (attention.py:96-97); the norm itself runs inside ltxb_qknorm_rope."""

    def __init__(self, weight: Tensor, eps: float) -> None:
        self.weight, self.eps = weight, eps


def _row_map(x: Tensor, rows: Tensor, index: Optional[Tensor]) -> Tuple[int, Optional[Tensor]]:
    """How token m of ``x`` (B, T, D) finds its row in a modulation tensor: (row divisor, row index)."""
    if index is not None:
        return 1, index
    B, T = x.shape[0], x.shape[1]
    n = rows.numel() // rows.shape[-1]
    if n == B * T:
        return 1, None
    if n == B:
        return max(T, 1), None
    raise ValueError(f"modulation tensor with {n} rows does not broadcast against {B}x{T} tokens")


class Attention:
    """attention.py:56-142.  Self-attention keeps to_q|to_k|to_v as ONE (3*inner, query_dim) matrix so
    the three projections are a single GEMM; cross-attention keeps to_k|to_v fused."""

    def __init__(self, query_dim: int, context_dim: Optional[int] = None, heads: int = 8, dim_head: int = 64,
                 norm_eps: float = 1e-6, rope_type: LTXRopeType = LTXRopeType.INTERLEAVED, device="cuda") -> None:
        self.rope_type, self.heads, self.dim_head = rope_type, heads, dim_head
        self.query_dim = query_dim
        self.is_self = context_dim is None
        self.context_dim = query_dim if context_dim is None else context_dim
        inner = self.inner_dim = heads * dim_head
        kw = dict(device=device)
        if self.is_self:
            self.qkv_weight = torch.zeros(3 * inner, query_dim, dtype=BF16, **kw)
            self.qkv_bias = torch.zeros(3 * inner, dtype=F32, **kw)
            w, b = self.qkv_weight, self.qkv_bias
            self.to_q = Linear(w[:inner], b[:inner])
            self.to_k = Linear(w[inner:2 * inner], b[inner:2 * inner])
            self.to_v = Linear(w[2 * inner:], b[2 * inner:])
        else:
            self.to_q = Linear(torch.zeros(inner, query_dim, dtype=BF16, **kw), torch.zeros(inner, dtype=F32, **kw))
            self.kv_weight = torch.zeros(2 * inner, self.context_dim, dtype=BF16, **kw)
            self.kv_bias = torch.zeros(2 * inner, dtype=F32, **kw)
            self.to_k = Linear(self.kv_weight[:inner], self.kv_bias[:inner])
            self.to_v = Linear(self.kv_weight[inner:], self.kv_bias[inner:])
        # q_norm | k_norm weights stored back to back: self-attention normalises and rotates q and k in ONE launch
        self.qk_norm_weight = torch.ones(2, inner, dtype=F32, **kw)
        self.q_norm = RMSNormWeight(self.qk_norm_weight[0], norm_eps)
        self.k_norm = RMSNormWeight(self.qk_norm_weight[1], norm_eps)
        self.to_out = Linear(torch.zeros(query_dim, inner, dtype=BF16, **kw), torch.zeros(query_dim, dtype=F32, **kw))

    # -- the pieces, exposed separately so the sequence-parallel path can put its all-to-all between them
    def project(self, ws: Workspace, tag: str, xq: Tensor, B: int, Tq: int, context: Optional[Tensor], Tk: int,
                pe, k_pe, kv_out: Optional[Tensor] = None, kv_ready: bool = False) -> Tuple[Tensor, Tensor, Tensor]:
        """Q/K/V projections + full-width q/k RMSNorm + split RoPE (attention.py:123-136).
        xq: bf16 [B*Tq, query_dim]; context: bf16 [B*Tk, context_dim] or None (self)."""
        inner, dev = self.inner_dim, xq.device
        if self.rope_type != LTXRopeType.SPLIT and pe is not None:
            raise _lib.LtxbError(f"rope_type {self.rope_type} has no sm_100a kernel; LTX-2 uses SPLIT")
        if context is None:
            if not self.is_self:
                raise ValueError("cross-attention module called without context")
            qkv = ws.get(tag + ".qkv", (B * Tq, 3 * inner), BF16, dev)
            ops.gemm(xq, self.qkv_weight, self.qkv_bias, qkv, const_w=True)
            q, k, v = qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:]
            Tk = Tq
        else:
            q = ws.get(tag + ".q", (B * Tq, inner), BF16, dev)
            ops.gemm(xq, self.to_q.weight, self.to_q.bias, q, const_w=True)
            kv = ws.get(tag + ".kv", (B * Tk, 2 * inner), BF16, dev) if kv_out is None else kv_out
            if not kv_ready:
                if self.is_self:  # self-attention weights applied to an explicit context (attention.py:124-126)
                    ops.gemm(context, self.qkv_weight[inner:], self.qkv_bias[inner:], kv, const_w=True)
                else:
                    ops.gemm(context, self.kv_weight, self.kv_bias, kv, const_w=True)
            k, v = kv[:, :inner], kv[:, inner:]
        H, dh = self.heads, self.dim_head
        if context is None and k_pe is None and self.k_norm.weight.data_ptr() == self.qk_norm_weight[1].data_ptr():
            cs = (None, None) if pe is None else pe
            ops.qknorm_rope_segments(qkv, 2, inner, B, Tq, H, dh, self.qk_norm_weight, self.q_norm.eps, cs[0], cs[1])
        elif pe is not None:
            kp = pe if k_pe is None else k_pe
            ops.qknorm_rope(q, B, Tq, H, dh, self.q_norm.weight, self.q_norm.eps, pe[0], pe[1])
            if not kv_ready:
                ops.qknorm_rope(k, B, Tk, H, dh, self.k_norm.weight, self.k_norm.eps, kp[0], kp[1])
        else:
            ops.qknorm_rope(q, B, Tq, H, dh, self.q_norm.weight, self.q_norm.eps)
            if not kv_ready:
                ops.qknorm_rope(k, B, Tk, H, dh, self.k_norm.weight, self.k_norm.eps)
        return q, k, v

    def sdpa(self, ws: Workspace, tag: str, q: Tensor, k: Tensor, v: Tensor, B: int, Tq: int, Tk: int,
             kv_bias: Optional[Tensor], heads: Optional[int] = None) -> Tensor:
        """attention.py:13-53 on [B*T, heads*dh] row-strided views."""
        H = self.heads if heads is None else heads
        o = ws.get(tag + ".o", (B * Tq, H * self.dim_head), BF16, q.device)
        ops.attention(q, k, v, o, B, Tq, Tk, H, self.dim_head, 1.0 / math.sqrt(self.dim_head), kv_bias)
        return o

    def fused(self, ws: Workspace, tag: str, xq: Tensor, B: int, Tq: int, resid: Tensor, *,
              context: Optional[Tensor] = None, Tk: int = 0, pe=None, k_pe=None, kv_bias: Optional[Tensor] = None,
              gate: Optional[Tensor] = None, gate_table: Optional[Tensor] = None, row_div: int = 1,
              row_index: Optional[Tensor] = None, seq_parallel=None, kv_out: Optional[Tensor] = None,
              kv_ready: bool = False, defer: bool = False) -> Optional[Tensor]:
        """resid (f32 [B*Tq, query_dim]) += to_out(attention(...)) * gate, in place — the to_out GEMM's
        epilogue carries bias, gate and residual add (transformer.py:254,257-261).
        ``defer=True``: the projection (bias included) is returned as bf16 instead and the CALLER adds it — the
        next sub-layer's norm kernel does, in the pass that reads the row anyway (ops.residual_rmsnorm_modulate)."""
        group_cols, peer_sync = 0, None
        if seq_parallel is not None and context is None and self.is_self:
            # rows are sharded across ranks: projections local, heads <-> sequence all-to-all around the attention
            o, group_cols = seq_parallel.self_attention(self, ws, tag, xq, B, Tq, pe)
            peer_sync = seq_parallel.take_pending_sync() if hasattr(seq_parallel, "take_pending_sync") else None
        else:
            q, k, v = self.project(ws, tag, xq, B, Tq, context, Tk, pe, k_pe, kv_out, kv_ready)
            o = self.sdpa(ws, tag, q, k, v, B, Tq, Tq if context is None else Tk, kv_bias)
        if defer:
            y = ws.get(tag + ".y", (B * Tq, self.query_dim), BF16, resid.device)
            ops.gemm(o, self.to_out.weight, self.to_out.bias, y, _lib.EPI_BIAS_BF16, a_group_cols=group_cols, const_w=True,
                     peer_sync=peer_sync)
            return y
        ops.gemm(o, self.to_out.weight, self.to_out.bias, resid, _lib.EPI_RESID_GATE_F32, resid=resid, gate=gate,
                 gate_table=gate_table, gate_row_div=row_div, gate_row_index=row_index, a_group_cols=group_cols, const_w=True,
                 peer_sync=peer_sync)
        return None

    def __call__(self, x: Tensor, context: Optional[Tensor] = None, mask: Optional[Tensor] = None, pe=None,
                 k_pe=None) -> Tensor:
        """Reference signature (attention.py:102-110): bf16 (B, T, query_dim) in -> bf16 out, un-fused."""
        if x.dtype != BF16:
            raise _lib.LtxbError("Attention.__call__ takes bf16 activations")
        ws = Workspace()
        B, Tq = x.shape[0], x.shape[1]
        ctx2 = None if context is None else context.reshape(-1, context.shape[-1])
        Tk = Tq if context is None else context.shape[1]
        q, k, v = self.project(ws, "a", x.reshape(B * Tq, -1), B, Tq, ctx2, Tk, pe, k_pe)
        o = self.sdpa(ws, "a", q, k, v, B, Tq, Tk, _kv_bias(mask, B, Tk))
        return self.to_out(o).view(B, Tq, self.query_dim)

    def named_parameters(self, prefix: str) -> Iterator[Tuple[str, Tensor]]:
        for n in ("to_q", "to_k", "to_v", "to_out"):
            lin = getattr(self, n)
            yield f"{prefix}.{n}.weight", lin.weight
            yield f"{prefix}.{n}.bias", lin.bias
        yield f"{prefix}.q_norm.weight", self.q_norm.weight
        yield f"{prefix}.k_norm.weight", self.k_norm.weight


def _kv_bias(mask: Optional[Tensor], B: int, Tk: int) -> Optional[Tensor]:
    """The kernels take a per-key additive bias f32 (B, Tk).  The reference's float masks are
    (B, 1, 1|Tq, Tk) (ltx.py:91-107; attention.py:36-42); a per-query mask has no kernel here."""
    if mask is None:
        return None
    m = mask
    if not m.dtype.is_floating_point:
        m = (m.to(F32) - 1.0) * 1e9
    if m.numel() != B * Tk:
        if m.numel() == Tk:
            m = m.reshape(1, Tk).expand(B, Tk)
        else:
            raise _lib.LtxbError(f"attention mask of shape {tuple(mask.shape)}: only per-key masks (B, Tk) are supported")
    return m.reshape(B, Tk).to(F32).contiguous()


class FeedForward:
    """feed_forward.py:17-40 — Linear D->4D, GELU(tanh), Linear 4D->D."""

    def __init__(self, dim: int, dim_out: Optional[int] = None, mult: int = 4, device="cuda") -> None:
        dim_out = dim if dim_out is None else dim_out
        inner = dim * mult
        self.proj_in = Linear(torch.zeros(inner, dim, dtype=BF16, device=device), torch.zeros(inner, dtype=F32, device=device))
        self.proj_out = Linear(torch.zeros(dim_out, inner, dtype=BF16, device=device), torch.zeros(dim_out, dtype=F32, device=device))

    def fused(self, ws: Workspace, tag: str, x: Tensor, resid: Tensor, gate, gate_table, row_div, row_index) -> None:
        h = ws.get(tag + ".h", (x.shape[0], self.proj_in.weight.shape[0]), BF16, x.device)
        ops.gemm(x, self.proj_in.weight, self.proj_in.bias, h, _lib.EPI_GELU_BF16, const_w=True)
        ops.gemm(h, self.proj_out.weight, self.proj_out.bias, resid, _lib.EPI_RESID_GATE_F32, resid=resid, gate=gate,
                 gate_table=gate_table, gate_row_div=row_div, gate_row_index=row_index, const_w=True)

    def __call__(self, x: Tensor) -> Tensor:
        h = self.proj_in(x, mode=_lib.EPI_GELU_BF16)
        return self.proj_out(h)

    def named_parameters(self, prefix: str) -> Iterator[Tuple[str, Tensor]]:
        for n in ("proj_in", "proj_out"):
            lin = getattr(self, n)
            yield f"{prefix}.{n}.weight", lin.weight
            yield f"{prefix}.{n}.bias", lin.bias


class BasicAVTransformerBlock:
    """transformer.py:38-361.  Video and audio streams, text cross-attention, audio<->video
    cross-attention, FFN — AdaLN shift/scale folded into the norm kernels, gate + residual folded
    into the GEMM epilogues."""

    def __init__(self, idx: int, video: Optional[TransformerConfig] = None, audio: Optional[TransformerConfig] = None,
                 rope_type: LTXRopeType = LTXRopeType.INTERLEAVED, norm_eps: float = 1e-6, device="cuda") -> None:
        self.idx, self.norm_eps = idx, norm_eps
        self.workspace = Workspace()
        a = dict(rope_type=rope_type, norm_eps=norm_eps, device=device)
        if video is not None:
            self.attn1 = Attention(query_dim=video.dim, heads=video.heads, dim_head=video.d_head, context_dim=None, **a)
            self.attn2 = Attention(query_dim=video.dim, context_dim=video.context_dim, heads=video.heads, dim_head=video.d_head, **a)
            self.ff = FeedForward(video.dim, dim_out=video.dim, device=device)
            self.scale_shift_table = torch.zeros(6, video.dim, dtype=F32, device=device)
        if audio is not None:
            self.audio_attn1 = Attention(query_dim=audio.dim, heads=audio.heads, dim_head=audio.d_head, context_dim=None, **a)
            self.audio_attn2 = Attention(query_dim=audio.dim, context_dim=audio.context_dim, heads=audio.heads, dim_head=audio.d_head, **a)
            self.audio_ff = FeedForward(audio.dim, dim_out=audio.dim, device=device)
            self.audio_scale_shift_table = torch.zeros(6, audio.dim, dtype=F32, device=device)
        if audio is not None and video is not None:
            self.audio_to_video_attn = Attention(query_dim=video.dim, context_dim=audio.dim, heads=audio.heads, dim_head=audio.d_head, **a)
            self.video_to_audio_attn = Attention(query_dim=audio.dim, context_dim=video.dim, heads=audio.heads, dim_head=audio.d_head, **a)
            self.scale_shift_table_a2v_ca_audio = torch.zeros(5, audio.dim, dtype=F32, device=device)
            self.scale_shift_table_a2v_ca_video = torch.zeros(5, video.dim, dtype=F32, device=device)

    # ------------------------------------------------------------------ stream pieces
    def _attn_pair(self, ws, tag, a: TransformerArgs, attn1: Attention, attn2: Attention, table: Tensor,
                   seq_parallel=None) -> Tensor:
        """x += attn1(rms(x)(1+scale)+shift, pe) * gate ; x += attn2(rms(x), context)   (transformer.py:247-261).
        The attn2 projection is RETURNED (bf16), not yet added: ``_add_pending`` or ``_ff`` adds it."""
        B, T, D = a.x.shape
        x2 = a.x.view(B * T, D)
        mod = a.timesteps.view(-1, a.timesteps.shape[-1])
        div, idx = _row_map(a.x, a.timesteps, a.timestep_index)
        nx = ws.get(tag + ".nx", (B * T, D), BF16, x2.device)
        # rows of the table / columns of the modulation: shift_msa, scale_msa, gate_msa (transformer.py:248)
        ops.rmsnorm_modulate(x2, nx, self.norm_eps, mod=mod, scale_off=D, shift_off=0, table_scale=table[1],
                             table_shift=table[0], row_div=div, row_index=idx)
        # the two out-projections keep the bf16 epilogue; their residual adds ride in the norm kernel that follows
        if DEFER_RESIDUAL:
            y1 = attn1.fused(ws, tag + ".attn1", nx, B, T, x2, pe=a.positional_embeddings, seq_parallel=seq_parallel, defer=True)
            ops.residual_rmsnorm_modulate(x2, y1, nx, self.norm_eps, mod=mod, gate_off=2 * D, table_gate=table[2],
                                          row_div=div, row_index=idx)  # x += attn1 * gate_msa ; nx = rms(x)
        else:
            attn1.fused(ws, tag + ".attn1", nx, B, T, x2, pe=a.positional_embeddings, gate=mod[:, 2 * D:3 * D],
                        gate_table=table[2], row_div=div, row_index=idx, seq_parallel=seq_parallel)
            ops.rmsnorm_modulate(x2, nx, self.norm_eps)
        Tc = a.context.shape[1]
        cache = a.context_cache
        if a.text_kv is not None:
            w = 2 * attn2.inner_dim
            kv_out, kv_ready = a.text_kv[:, self.idx * w:(self.idx + 1) * w], True
        else:
            kv_out = None if cache is None else cache.entry(self.idx, (a.context.shape[0] * Tc, 2 * attn2.inner_dim), x2.device)
            kv_ready = cache is not None and cache.valid
        return attn2.fused(ws, tag + ".attn2", nx, B, T, x2, context=a.context.reshape(-1, a.context.shape[-1]), Tk=Tc,
                           kv_bias=a.context_mask, kv_out=kv_out, kv_ready=kv_ready, defer=DEFER_RESIDUAL)

    @staticmethod
    def _add_pending(a: TransformerArgs, y: Optional[Tensor]) -> None:
        """x += y for a deferred (ungated) projection that no norm kernel will pick up before x is read."""
        if y is not None:
            ops.gate_residual(a.x.view(-1, a.x.shape[-1]), y)

    def _ff(self, ws, tag, a: TransformerArgs, ff: FeedForward, table: Tensor, pending: Optional[Tensor] = None) -> None:
        """[x += pending ;] x += ff(rms(x)(1+scale)+shift) * gate   (transformer.py:342-355)"""
        B, T, D = a.x.shape
        x2 = a.x.view(B * T, D)
        mod = a.timesteps.view(-1, a.timesteps.shape[-1])
        div, idx = _row_map(a.x, a.timesteps, a.timestep_index)
        nx = ws.get(tag + ".nx", (B * T, D), BF16, x2.device)
        if pending is not None:
            ops.residual_rmsnorm_modulate(x2, pending, nx, self.norm_eps, mod=mod, scale_off=4 * D, shift_off=3 * D,
                                          table_scale=table[4], table_shift=table[3], row_div=div, row_index=idx)
        else:
            ops.rmsnorm_modulate(x2, nx, self.norm_eps, mod=mod, scale_off=4 * D, shift_off=3 * D, table_scale=table[4],
                                 table_shift=table[3], row_div=div, row_index=idx)
        ff.fused(ws, tag + ".ff", nx, x2, mod[:, 5 * D:6 * D], table[5], div, idx)

    def _cross_av(self, ws, v: TransformerArgs, a: TransformerArgs, seq_parallel=None) -> None:
        """Audio<->video cross-attention (transformer.py:281-339).  Both directions read the streams as
        they were BEFORE either update, so all four modulated inputs are produced first."""
        eps = self.norm_eps
        Bv, Tv, Dv = v.x.shape
        Ba, Ta, Da = a.x.shape
        vx, ax = v.x.view(Bv * Tv, Dv), a.x.view(Ba * Ta, Da)
        tv, ta = self.scale_shift_table_a2v_ca_video, self.scale_shift_table_a2v_ca_audio
        vss = v.cross_scale_shift_timestep.view(-1, 4 * Dv)
        ass = a.cross_scale_shift_timestep.view(-1, 4 * Da)
        vg = v.cross_gate_timestep.view(-1, Dv)
        ag = a.cross_gate_timestep.view(-1, Da)
        vdiv, vidx = _row_map(v.x, v.cross_scale_shift_timestep, v.timestep_index)
        adiv, aidx = _row_map(a.x, a.cross_scale_shift_timestep, a.timestep_index)
        dev = vx.device
        v_a2v = ws.get("av.v_a2v", (Bv * Tv, Dv), BF16, dev)
        a_a2v = ws.get("av.a_a2v", (Ba * Ta, Da), BF16, dev)
        a_v2a = ws.get("av.a_v2a", (Ba * Ta, Da), BF16, dev)
        v_v2a = ws.get("av.v_v2a", (Bv * Tv, Dv), BF16, dev)
        # columns / rows 0..3 = scale_a2v, shift_a2v, scale_v2a, shift_v2a (transformer.py:179-219)
        ops.rmsnorm_modulate(vx, v_a2v, eps, mod=vss, scale_off=0, shift_off=Dv, table_scale=tv[0], table_shift=tv[1], row_div=vdiv, row_index=vidx)
        ops.rmsnorm_modulate(ax, a_a2v, eps, mod=ass, scale_off=0, shift_off=Da, table_scale=ta[0], table_shift=ta[1], row_div=adiv, row_index=aidx)
        ops.rmsnorm_modulate(ax, a_v2a, eps, mod=ass, scale_off=2 * Da, shift_off=3 * Da, table_scale=ta[2], table_shift=ta[3], row_div=adiv, row_index=aidx)
        ops.rmsnorm_modulate(vx, v_v2a, eps, mod=vss, scale_off=2 * Dv, shift_off=3 * Dv, table_scale=tv[2], table_shift=tv[3], row_div=vdiv, row_index=vidx)
        # a2v: Q from video, K/V from audio, gated by the video table's row 4
        self.audio_to_video_attn.fused(ws, "av.a2v", v_a2v, Bv, Tv, vx, context=a_a2v, Tk=Ta,
                                       pe=v.cross_positional_embeddings, k_pe=a.cross_positional_embeddings,
                                       gate=vg, gate_table=tv[4], row_div=vdiv, row_index=vidx)
        # v2a: Q from audio, K/V from video, gated by the audio table's row 4
        if seq_parallel is not None:
            seq_parallel.video_to_audio(self.video_to_audio_attn, ws, a_v2a, v_v2a, Ba, Ta, Tv, ax, a, v, ag, ta[4], adiv, aidx)
        else:
            self.video_to_audio_attn.fused(ws, "av.v2a", a_v2a, Ba, Ta, ax, context=v_v2a, Tk=Tv,
                                           pe=a.cross_positional_embeddings, k_pe=v.cross_positional_embeddings,
                                           gate=ag, gate_table=ta[4], row_div=adiv, row_index=aidx)

    # ------------------------------------------------------------------ the block
    def __call__(self, video: Optional[TransformerArgs] = None, audio: Optional[TransformerArgs] = None, *,
                 inplace: bool = False, workspace: Optional[Workspace] = None, seq_parallel=None,
                 ) -> Tuple[Optional[TransformerArgs], Optional[TransformerArgs]]:
        """transformer.py:221-361.  With ``inplace=False`` (the reference's value semantics) the residual
        streams are copied first; the model's own block loop passes ``inplace=True``."""
        ws = self.workspace if workspace is None else workspace
        run_vx = video is not None and video.enabled and video.x.numel() > 0
        run_ax = audio is not None and audio.enabled and audio.x.numel() > 0
        if not inplace:
            if run_vx:
                video = replace(video, x=video.x.clone())
            if run_ax:
                audio = replace(audio, x=audio.x.clone())
        yv = ya = None  # text-cross-attention projections still to be added to the streams
        if run_vx:
            yv = self._attn_pair(ws, "v", video, self.attn1, self.attn2, self.scale_shift_table, seq_parallel)
        if run_ax:
            ya = self._attn_pair(ws, "a", audio, self.audio_attn1, self.audio_attn2, self.audio_scale_shift_table)
        if run_vx and run_ax:
            # the audio<->video cross-attention reads both streams: they must be complete first
            self._add_pending(video, yv)
            self._add_pending(audio, ya)
            yv = ya = None
            self._cross_av(ws, video, audio, seq_parallel)
        if run_vx:
            self._ff(ws, "v", video, self.ff, self.scale_shift_table, yv)
        if run_ax:
            self._ff(ws, "a", audio, self.audio_ff, self.audio_scale_shift_table, ya)
        return video, audio

    def named_parameters(self, prefix: str) -> Iterator[Tuple[str, Tensor]]:
        for name in ("attn1", "attn2", "ff", "audio_attn1", "audio_attn2", "audio_ff", "audio_to_video_attn",
                     "video_to_audio_attn"):
            if hasattr(self, name):
                yield from getattr(self, name).named_parameters(f"{prefix}.{name}")
        for name in ("scale_shift_table", "audio_scale_shift_table", "scale_shift_table_a2v_ca_audio",
                     "scale_shift_table_a2v_ca_video"):
            if hasattr(self, name):
                yield f"{prefix}.{name}", getattr(self, name)
