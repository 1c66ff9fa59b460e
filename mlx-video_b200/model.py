"""LTX-2 DiT model (velocity prediction) on sm_100a kernels — the drop-in boundary.

Mirror of the reference's ``mlx_video/models/ltx/ltx.py``: ``LTXModel`` (:250-506; preprocessors :33-247,
output head :432-457), ``X0Model`` (:888-906), ``adaln.py:9-138`` and ``text_projection.py:5-26``.
Same constructor (``LTXModel(config)``), same call signature
(``model(video: Modality | None, audio: Modality | None) -> (velocity_video, velocity_audio)``), same
attributes the samplers read (``inner_dim``, ``positional_embedding_theta`` ..., generate.py:614-637), same
exceptions (``ValueError`` for a modality the model type lacks, ltx.py:466-469; ``ValueError`` for
parameters missing at a strict load, ltx.py:874-881).  All arithmetic runs in ``libltxb.so``; torch
supplies device memory and the stream.  There is no CPU path: CPU tensors raise ``LtxbError``.
"""
from __future__ import annotations

import math
import os
from dataclasses import replace
from pathlib import Path
from typing import Dict, Iterator, List, Optional, Tuple, Union

import torch

from . import _lib, ops
from ._lib import LtxbError
from .config import LTXModelConfig, LTXModelType, LTXRopeType
from .rope import precompute_freqs_cis
from .transformer import (BF16, F32, BasicAVTransformerBlock, ContextCache, Linear, Modality, TransformerArgs, Workspace,
                          _kv_bias)

Tensor = torch.Tensor


def _new_linear(n_out: int, n_in: int, device) -> Linear:
    return Linear(torch.zeros(n_out, n_in, dtype=BF16, device=device), torch.zeros(n_out, dtype=F32, device=device))


def _lin_params(prefix: str, lin: Linear) -> Iterator[Tuple[str, Tensor]]:
    yield prefix + ".weight", lin.weight
    yield prefix + ".bias", lin.bias


class _Holder:
    """Attribute bag so parameter paths read like the reference's module tree."""


class AdaLayerNormSingle:
    """adaln.py:9-47 — sinusoid(256, cos|sin) -> Linear -> SiLU -> Linear (= embedded_timestep) -> SiLU
    -> Linear(dim -> k*dim).  Runs on one row per DISTINCT timestep."""

    def __init__(self, embedding_dim: int, embedding_coefficient: int = 6, device="cuda") -> None:
        self.embedding_dim, self.embedding_coefficient = embedding_dim, embedding_coefficient
        self.emb = _Holder()
        self.emb.timestep_embedder = _Holder()
        self.emb.timestep_embedder.linear1 = _new_linear(embedding_dim, 256, device)
        self.emb.timestep_embedder.linear2 = _new_linear(embedding_dim, embedding_dim, device)
        self.linear = _new_linear(embedding_coefficient * embedding_dim, embedding_dim, device)

    def __call__(self, timestep: Tensor, scale: float = 1.0, hidden_dtype=None) -> Tuple[Tensor, Tensor]:
        """timestep f32 [n] (already multiplied by timestep_scale_multiplier unless ``scale`` says so) ->
        (modulation f32 [n, k*dim], embedded_timestep f32 [n, dim])."""
        n, dim, dev = timestep.numel(), self.embedding_dim, timestep.device
        te = self.emb.timestep_embedder
        feat = torch.empty(n, 256, dtype=BF16, device=dev)
        ops.timestep_embed(timestep.reshape(-1), scale, 256, feat)
        h = te.linear1(feat, mode=_lib.EPI_SILU_BF16)
        e = te.linear2(h)
        se = torch.empty_like(e)
        ops.silu_bf16(e, se)
        mod = self.linear(se, mode=_lib.EPI_BIAS_F32)
        e32 = torch.empty(n, dim, dtype=F32, device=dev)
        ops.cast_bf16_to_f32(e, e32)
        return mod, e32

    def named_parameters(self, prefix: str) -> Iterator[Tuple[str, Tensor]]:
        yield from _lin_params(prefix + ".emb.timestep_embedder.linear1", self.emb.timestep_embedder.linear1)
        yield from _lin_params(prefix + ".emb.timestep_embedder.linear2", self.emb.timestep_embedder.linear2)
        yield from _lin_params(prefix + ".linear", self.linear)


class PixArtAlphaTextProjection:
    """text_projection.py:5-26 — Linear, GELU(tanh), Linear."""

    def __init__(self, in_features: int, hidden_size: int, out_features: Optional[int] = None, device="cuda") -> None:
        out_features = hidden_size if out_features is None else out_features
        self.linear1 = _new_linear(hidden_size, in_features, device)
        self.linear2 = _new_linear(out_features, hidden_size, device)

    def __call__(self, caption: Tensor) -> Tensor:
        return self.linear2(self.linear1(caption, mode=_lib.EPI_GELU_BF16))

    def named_parameters(self, prefix: str) -> Iterator[Tuple[str, Tensor]]:
        yield from _lin_params(prefix + ".linear1", self.linear1)
        yield from _lin_params(prefix + ".linear2", self.linear2)


def _as_bf16(t: Tensor) -> Tensor:
    """bf16 contiguous copy of an activation through ltxb_cast (no torch arithmetic on the path)."""
    if not t.is_cuda:
        raise LtxbError("LTXModel inputs must be CUDA tensors; there is no CPU fallback on this path")
    t = t.contiguous()
    if t.dtype == BF16:
        return t
    if t.dtype != F32:
        raise LtxbError(f"unsupported activation dtype {t.dtype} (bf16 or f32)")
    out = torch.empty(t.shape, dtype=BF16, device=t.device)
    ops.cast_f32_to_bf16(t, out)
    return out


class LTXModel:
    """ltx.py:250-506.  ``dedupe_timesteps``: group equal per-token timesteps on the device so that
    AdaLN-single runs over the distinct values only (at most ``timestep_capacity``); False reproduces
    the reference's one-row-per-token evaluation literally."""

    def __init__(self, config: LTXModelConfig, device: Union[str, torch.device, None] = None,
                 dedupe_timesteps: bool = True, timestep_capacity: int = 128, cache_context: bool = False,
                 cuda_graphs: bool = False, context_cache_slots: int = 2) -> None:
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        if device is None or torch.device(device).type != "cuda":
            raise LtxbError("LTXModel needs a CUDA device (sm_100a); there is no CPU fallback on this path")
        self.device = torch.device(device)
        self.config = config
        self.model_type = config.model_type
        self.use_middle_indices_grid = config.use_middle_indices_grid
        self.rope_type = config.rope_type
        self.timestep_scale_multiplier = config.timestep_scale_multiplier
        self.positional_embedding_theta = config.positional_embedding_theta
        self.dedupe_timesteps, self.timestep_capacity = dedupe_timesteps, timestep_capacity
        # cache_context: keep the caption projection and every block's text K/V while the SAME context tensor keeps
        # coming in (constant across a denoise loop; the reference recomputes them every step — off by default so
        # the default forward executes the reference's work).  cuda_graphs: replay one captured graph per input
        # signature instead of ~800 launches per forward.
        # context_cache_slots: how many DIFFERENT contexts keep their projections at once per modality (2 = the cond /
        # uncond prompts the two-pass CFG loops alternate between, generate.py:1258-1283); least recently used goes first.
        self.cache_context, self._context_caches, self._context_keys = cache_context, {}, {}
        self.context_cache_slots = max(1, int(context_cache_slots))
        self._graphs = {} if cuda_graphs else None
        self.workspace = Workspace()
        self.seq_parallel = None  # set by parallel.UlyssesGroup.attach()
        self._group_counts: List[Tensor] = []
        dev = self.device
        if config.model_type.is_video_enabled():
            self.positional_embedding_max_pos = config.positional_embedding_max_pos
            self.num_attention_heads = config.num_attention_heads
            self.inner_dim = config.inner_dim
            self.patchify_proj = _new_linear(self.inner_dim, config.in_channels, dev)
            self.adaln_single = AdaLayerNormSingle(self.inner_dim, device=dev)
            self.caption_projection = PixArtAlphaTextProjection(config.caption_channels, self.inner_dim, device=dev)
            self.scale_shift_table = torch.zeros(2, self.inner_dim, dtype=F32, device=dev)
            self.proj_out = _new_linear(config.out_channels, self.inner_dim, dev)
        if config.model_type.is_audio_enabled():
            self.audio_positional_embedding_max_pos = config.audio_positional_embedding_max_pos
            self.audio_num_attention_heads = config.audio_num_attention_heads
            self.audio_inner_dim = config.audio_inner_dim
            self.audio_patchify_proj = _new_linear(self.audio_inner_dim, config.audio_in_channels, dev)
            self.audio_adaln_single = AdaLayerNormSingle(self.audio_inner_dim, device=dev)
            self.audio_caption_projection = PixArtAlphaTextProjection(config.audio_caption_channels, self.audio_inner_dim, device=dev)
            self.audio_scale_shift_table = torch.zeros(2, self.audio_inner_dim, dtype=F32, device=dev)
            self.audio_proj_out = _new_linear(config.audio_out_channels, self.audio_inner_dim, dev)
        self._av = config.model_type.is_video_enabled() and config.model_type.is_audio_enabled()
        if self._av:
            self.cross_pe_max_pos = max(config.positional_embedding_max_pos[0], config.audio_positional_embedding_max_pos[0])
            self.av_ca_timestep_scale_multiplier = config.av_ca_timestep_scale_multiplier
            self.audio_cross_attention_dim = config.audio_cross_attention_dim
            self.av_ca_video_scale_shift_adaln_single = AdaLayerNormSingle(self.inner_dim, 4, device=dev)
            self.av_ca_audio_scale_shift_adaln_single = AdaLayerNormSingle(self.audio_inner_dim, 4, device=dev)
            self.av_ca_a2v_gate_adaln_single = AdaLayerNormSingle(self.inner_dim, 1, device=dev)
            self.av_ca_v2a_gate_adaln_single = AdaLayerNormSingle(self.audio_inner_dim, 1, device=dev)
        vcfg, acfg = config.get_video_config(), config.get_audio_config()
        self.transformer_blocks: Dict[int, BasicAVTransformerBlock] = {
            i: BasicAVTransformerBlock(idx=i, video=vcfg, audio=acfg, rope_type=config.rope_type, norm_eps=config.norm_eps, device=dev)
            for i in range(config.num_layers)
        }
        # memory layout for the hoisted text K/V projection: every block's attn2 to_k|to_v weight in ONE
        # (L*2*inner, context_dim) matrix, k_norm weights in one (L, inner) table; the blocks hold views
        self._text_kv = {}
        stack = os.environ.get("LTXB_STACK_TEXT_KV", "1") != "0"  # 0: per-block K/V GEMMs (A/B runs)
        if stack and config.model_type.is_video_enabled():
            self._stack_text_kv("", "attn2")
        if stack and config.model_type.is_audio_enabled():
            self._stack_text_kv("audio_", "audio_attn2")

    def _stack_text_kv(self, pre: str, name: str) -> None:
        blocks = [getattr(b, name) for b in self.transformer_blocks.values()]
        L, inner, cdim = len(blocks), blocks[0].inner_dim, blocks[0].context_dim
        w = torch.zeros(L * 2 * inner, cdim, dtype=BF16, device=self.device)
        b = torch.zeros(L * 2 * inner, dtype=F32, device=self.device)
        kn = torch.ones(L, inner, dtype=F32, device=self.device)
        for l, att in enumerate(blocks):
            r = slice(l * 2 * inner, (l + 1) * 2 * inner)
            att.kv_weight, att.kv_bias = w[r], b[r]
            att.to_k = Linear(att.kv_weight[:inner], att.kv_bias[:inner])
            att.to_v = Linear(att.kv_weight[inner:], att.kv_bias[inner:])
            att.k_norm.weight = kn[l]
        self._text_kv[pre] = (w, b, kn, blocks[0].heads, blocks[0].dim_head, blocks[0].k_norm.eps)

    def _project_text_kv(self, pre: str, ctx: Tensor, cache: Optional[ContextCache]) -> Tensor:
        """K|V of the text cross-attention of every block (attention.py:124-130 for attn2 of all 48 blocks): the
        context is the same for all of them, so it is one (B*Tc, D) x (D, L*2*inner) GEMM and one norm launch."""
        w, b, kn, H, dh, eps = self._text_kv[pre]
        Bc, Tc = ctx.shape[0], ctx.shape[1]
        shape = (Bc * Tc, w.shape[0])
        kv = self.workspace.get(pre + "text_kv", shape, BF16, self.device) if cache is None else cache.stacked(shape, self.device)
        if cache is None or not cache.valid:
            ops.gemm(ctx.reshape(Bc * Tc, -1), w, b, kv, const_w=True)
            ops.qknorm_rope_segments(kv, kn.shape[0], 2 * H * dh, Bc, Tc, H, dh, kn, eps)
        return kv

    # ------------------------------------------------------------------ parameters
    def named_parameters(self) -> Iterator[Tuple[str, Tensor]]:
        """(name, tensor) with the reference's sanitised state-dict names (ltx.py:508-533)."""
        prefixes = ([""] if self.model_type.is_video_enabled() else []) + (["audio_"] if self.model_type.is_audio_enabled() else [])
        for pre in prefixes:
            yield from _lin_params(pre + "patchify_proj", getattr(self, pre + "patchify_proj"))
            yield from getattr(self, pre + "adaln_single").named_parameters(pre + "adaln_single")
            yield from getattr(self, pre + "caption_projection").named_parameters(pre + "caption_projection")
            yield pre + "scale_shift_table", getattr(self, pre + "scale_shift_table")
            yield from _lin_params(pre + "proj_out", getattr(self, pre + "proj_out"))
        if self._av:
            for n in ("av_ca_video_scale_shift_adaln_single", "av_ca_audio_scale_shift_adaln_single",
                      "av_ca_a2v_gate_adaln_single", "av_ca_v2a_gate_adaln_single"):
                yield from getattr(self, n).named_parameters(n)
        for i, blk in self.transformer_blocks.items():
            yield from blk.named_parameters(f"transformer_blocks.{i}")

    def parameters(self) -> Dict[str, Tensor]:
        return dict(self.named_parameters())

    def num_parameters(self) -> int:
        return sum(t.numel() for _, t in self.named_parameters())

    def load_weights(self, weights: Dict[str, Tensor], strict: bool = True, quant_meta: Optional[dict] = None,
                     keep_packed: bool = False) -> None:
        """Copy a state dict (reference names) into the device layout.  Strict: every model parameter
        must be present (ValueError otherwise, ltx.py:874-881) and unknown keys are an error.
        MLX affine-quantised linears (``X.weight`` uint32 beside ``X.scales`` / ``X.biases``, the tensors
        ``nn.QuantizedLinear`` holds after ltx.py:641-725) are expanded to bf16 on the device at this point
        (``ltxb_dequant_affine_bf16``); group size and bits follow from the tensor shapes and must agree with
        ``quant_meta`` (the checkpoint's quantization.json, ltx.py:649-668) when that is given.
        keep_packed: the packed words, scales and biases also stay on the device (packed.py) and few-row products
        (M <= 256: sequence-parallel shards, the audio stream, AdaLN rows) stream THEM instead of the expansion
        (``ltxb_gemm_qw_bf16``, bit-identical results, a quarter / half of the weight bytes)."""
        from .checkpoint import quant_layout
        from .packed import REGISTRY as packed_registry

        params = self.parameters()
        quantised = {k[: -len(".scales")] for k in weights if k.endswith(".scales")}
        aux = {f"{b}{suffix}" for b in quantised for suffix in (".scales", ".biases")}
        missing = [k for k in params if k not in weights]
        extra = [k for k in weights if k not in params and not (k in aux and k.rsplit(".", 1)[0] + ".weight" in params)]
        if strict and missing:
            raise ValueError(f"Missing {len(missing)} parameters: {missing[:8]}{'...' if len(missing) > 8 else ''}")
        if strict and extra:
            raise ValueError(f"Unexpected {len(extra)} parameters: {extra[:8]}{'...' if len(extra) > 8 else ''}")
        for name, dst in params.items():
            src = weights.get(name)
            if src is None:
                continue
            base = name[: -len(".weight")] if name.endswith(".weight") else None
            if base in quantised:
                scales, biases = weights[f"{base}.scales"], weights.get(f"{base}.biases")
                if biases is None:
                    raise ValueError(f"quantised linear {base} has .scales but no .biases")
                group_size, bits = quant_layout(name, tuple(src.shape), tuple(scales.shape), tuple(dst.shape), quant_meta)
                aux_dtype = BF16 if scales.dtype == BF16 else F32  # f16 scales are widened (plumbing), bf16 / f32 are read as stored
                dev = dst.device
                src_d, scales_d, biases_d = src.to(dev).contiguous(), scales.to(dev, aux_dtype).contiguous(), biases.to(dev, aux_dtype).contiguous()
                ops.dequant_affine(src_d, scales_d, biases_d, dst, group_size, bits)
                if keep_packed:
                    packed_registry.register(dst, src_d, scales_d, biases_d, group_size, bits)
                continue
            if src.dtype in (torch.uint32, torch.int32, torch.uint8):
                raise ValueError(f"{name} is an integer tensor without .scales / .biases siblings")
            if tuple(src.shape) != tuple(dst.shape):
                raise ValueError(f"shape mismatch for {name}: checkpoint {tuple(src.shape)} vs model {tuple(dst.shape)}")
            dst.copy_(src.to(device=dst.device, dtype=dst.dtype))  # plumbing: H2D copy + storage cast
        if keep_packed:
            packed_registry.merge_adjacent()
        self.clear_caches()  # projected-context caches and captured graphs belong to the old weights

    def sanitize(self, weights: Dict[str, Tensor]) -> Dict[str, Tensor]:
        """ltx.py:508-533: upstream (``model.diffusion_model.*``) names -> parameter names; every tensor without that
        prefix, and the embeddings connectors, are dropped.  Values are passed through untouched."""
        from .checkpoint import PREFIX, sanitize_key

        out: Dict[str, Tensor] = {}
        for key, value in weights.items():
            name = sanitize_key(key) if key.startswith(PREFIX) else None
            if name is not None:
                out[name] = value
        return out

    def init_random(self, seed: int = 0, table_std: float = 0.02) -> "LTXModel":
        """Random-init weights of the reference architecture, generated on the device (there is no
        checkpoint offline): Linear U(+-1/sqrt(in)) like nn.Linear, norm weights 1 + 0.1 N(0,1), tables
        0.02 N(0,1) (non-zero so AdaLN shift/scale/gate are exercised) — SURVEY.md §8c."""
        g = torch.Generator(device=self.device).manual_seed(seed)
        params = self.parameters()
        for name, p in params.items():
            if name.endswith("scale_shift_table") or "scale_shift_table_a2v" in name:
                p.copy_(torch.randn(p.shape, generator=g, device=p.device) * table_std)
            elif name.endswith("_norm.weight"):
                p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g, device=p.device))
            elif name.endswith(".weight"):
                k = 1.0 / math.sqrt(p.shape[1])
                for r0 in range(0, p.shape[0], 8192):  # bounded fp32 temporaries
                    rows = p[r0:r0 + 8192]
                    rows.copy_((torch.rand(rows.shape, generator=g, device=p.device) * 2 - 1) * k)
            else:  # bias: same bound as its weight, U(+-1/sqrt(fan_in))
                fan_in = params[name[:-len("bias")] + "weight"].shape[1]
                p.copy_((torch.rand(p.shape, generator=g, device=p.device) * 2 - 1) / math.sqrt(fan_in))
        return self

    @classmethod
    def from_pretrained(cls, model_path, config: LTXModelConfig, strict: bool = True,
                        weights_override: Optional[Dict[str, Tensor]] = None, device=None, keep_packed: bool = False) -> "LTXModel":
        """ltx.py:535-885: reads safetensors file(s), renames upstream keys, rounds fp32 tensors to bf16 values
        (ltx.py:613-615), expands MLX-quantised linears (ltx.py:641-725), IGNORES tensors the configured model does
        not have (audio weights beside a video-only config, ltx.py:739-740) and, when ``strict``, raises ValueError
        for every parameter the file(s) did not supply (ltx.py:874-881).  ``weights_override``: tensors already in
        memory (unified weights / LoRA merges, ltx.py:617-623); those must be the complete set when ``strict``."""
        from .checkpoint import checkpoint_files, load_transformer_weights, read_quantization_meta, sanitize_state_dict

        model = cls(config, device=device)
        expected = set(model.parameters())
        expected |= {k[: -len("weight")] + suffix for k in expected if k.endswith(".weight") for suffix in ("scales", "biases")}
        if weights_override is not None:
            # ltx.py:828-842: overrides go through the same key filter as file tensors, so a unified audio+video
            # state dict (the LoRA flow: apply_lora_to_weights -> from_pretrained(weights_override=merged)) loads into
            # a video-only or audio-only config; only MISSING parameters raise
            given = {k: v for k, v in sanitize_state_dict(dict(weights_override)).items() if k in expected}
            meta = None
            if any(k.endswith(".scales") for k in given) and model_path is not None:
                try:
                    meta = read_quantization_meta(checkpoint_files(model_path)[0])
                except (OSError, ValueError, IndexError):
                    meta = None
            model.load_weights(given, strict=strict, quant_meta=meta, keep_packed=keep_packed)
            return model
        weights = load_transformer_weights(model_path, config, expected=expected)
        meta = read_quantization_meta(checkpoint_files(model_path)[0]) if any(k.endswith(".scales") for k in weights) else None
        try:
            model.load_weights(weights, strict=strict, quant_meta=meta, keep_packed=keep_packed)
        except ValueError as e:
            if str(e).startswith("Missing "):
                raise ValueError(str(e).replace(" parameters:", " parameters after load (sample:", 1) + ").") from None
            raise
        return model

    # ------------------------------------------------------------------ preprocessors (ltx.py:33-247)
    def _timestep_rows(self, timesteps: Tensor, B: int, T: int) -> Tuple[Tensor, Optional[Tensor], tuple]:
        """-> (distinct-or-literal sigma values f32 [R], per-token row index or None, view shape (b, t))."""
        ts = timesteps
        if not ts.is_cuda:
            raise LtxbError("Modality.timesteps must be a CUDA tensor")
        ts = ts.to(F32).contiguous()  # plumbing cast of a (B, T) scalar field
        if ts.numel() == B:  # (B,) or (B, 1): one row per batch element
            return ts.reshape(B), None, (B, 1)
        if ts.numel() != B * T:
            raise ValueError(f"timesteps of shape {tuple(timesteps.shape)} do not match {B}x{T} tokens")
        if not self.dedupe_timesteps:
            return ts.reshape(B * T), None, (B, T)
        values, index, count = ops.timestep_groups(ts.reshape(-1), self.timestep_capacity)
        self._group_counts.append(count)
        if len(self._group_counts) > 64 and not torch.cuda.is_current_stream_capturing():
            # long-lived processes that never call check_timestep_groups(): fold the history into one device scalar
            self._group_counts = [torch.stack(self._group_counts).max().reshape(1)]
        return values, index, (1, self.timestep_capacity)

    def check_timestep_groups(self) -> None:
        """Synchronising check that no forward since the last call overflowed ``timestep_capacity`` (an
        overflow also poisons that forward's output with NaN, so it cannot pass silently)."""
        counts, self._group_counts = self._group_counts, []
        if counts and int(torch.stack(counts).max().item()) > self.timestep_capacity:
            raise LtxbError(f"more than {self.timestep_capacity} distinct timesteps in one forward; "
                            "raise timestep_capacity or pass dedupe_timesteps=False")

    def _prepare(self, m: Modality, audio: bool) -> TransformerArgs:
        c = self.config
        pre = "audio_" if audio else ""
        inner = self.audio_inner_dim if audio else self.inner_dim
        heads = self.audio_num_attention_heads if audio else self.num_attention_heads
        max_pos = self.audio_positional_embedding_max_pos if audio else self.positional_embedding_max_pos
        lat = _as_bf16(m.latent)
        B, T = lat.shape[0], lat.shape[1]
        x = getattr(self, pre + "patchify_proj")(lat, mode=_lib.EPI_BIAS_F32)  # fresh f32 (B, T, D): the residual stream
        values, index, (rb, rt) = self._timestep_rows(m.timesteps, B, T)
        scale = float(self.timestep_scale_multiplier)
        mod, emb = getattr(self, pre + "adaln_single")(values, scale)
        cache = self._context_cache_for(pre, m.context) if self.cache_context else None
        proj = getattr(self, pre + "caption_projection")
        if cache is None:
            ctx = proj(_as_bf16(m.context)).view(m.context.shape[0], -1, inner)
        else:
            ctx = cache.context_buffer((m.context.shape[0], m.context.shape[1], inner), self.device)
            if not cache.valid:
                proj.linear2(proj.linear1(_as_bf16(m.context), mode=_lib.EPI_GELU_BF16), out=ctx)
        text_kv = self._project_text_kv(pre, ctx, cache) if pre in self._text_kv else None
        mask = _kv_bias(m.context_mask, B, ctx.shape[1])
        pe = m.positional_embeddings
        if pe is None:
            pe = precompute_freqs_cis(m.positions.to(self.device), inner, self.positional_embedding_theta, max_pos,
                                      self.use_middle_indices_grid, heads, self.rope_type, c.double_precision_rope)
        args = TransformerArgs(x=x, context=ctx, context_mask=mask, timesteps=mod.view(rb, rt, -1),
                               embedded_timestep=emb.view(rb, rt, -1), positional_embeddings=pe, enabled=m.enabled,
                               timestep_index=index, context_cache=cache, text_kv=text_kv)
        if self._av:  # ltx.py:201-247
            cross_pe = precompute_freqs_cis(m.positions[:, 0:1].to(self.device), self.audio_cross_attention_dim,
                                            self.positional_embedding_theta, [self.cross_pe_max_pos], True, heads,
                                            self.rope_type, c.double_precision_rope)
            ss_mod = self.av_ca_audio_scale_shift_adaln_single if audio else self.av_ca_video_scale_shift_adaln_single
            g_mod = self.av_ca_v2a_gate_adaln_single if audio else self.av_ca_a2v_gate_adaln_single
            factor = self.av_ca_timestep_scale_multiplier / self.timestep_scale_multiplier
            ss, _ = ss_mod(values, scale)
            gt, _ = g_mod(values, scale * factor)
            args = replace(args, cross_positional_embeddings=cross_pe, cross_scale_shift_timestep=ss.view(rb, rt, -1),
                           cross_gate_timestep=gt.view(rb, rt, -1))
        return args

    def _process_output(self, table: Tensor, proj_out: Linear, a: TransformerArgs, out_dtype) -> Tensor:
        """ltx.py:432-457 — LayerNorm(no affine) * (1 + scale) + shift (table row 0 = shift, row 1 = scale,
        plus embedded_timestep), then Linear D -> out_channels."""
        from .transformer import _row_map

        B, T, D = a.x.shape
        div, idx = _row_map(a.x, a.embedded_timestep, a.timestep_index)
        nx = self.workspace.get("out.nx", (B * T, D), BF16, a.x.device)
        ops.layernorm_modulate(a.x.view(B * T, D), nx, self.config.norm_eps, a.embedded_timestep.view(-1, D), table[1],
                               table[0], div, idx)
        if out_dtype == BF16:
            return proj_out(nx).view(B, T, -1)
        return proj_out(nx, mode=_lib.EPI_BIAS_F32).view(B, T, -1)

    # ------------------------------------------------------------------ forward (ltx.py:459-506)
    def prepare(self, video: Optional[Modality], audio: Optional[Modality]):
        if not self.model_type.is_video_enabled() and video is not None:
            raise ValueError("Video is not enabled for this model")
        if not self.model_type.is_audio_enabled() and audio is not None:
            raise ValueError("Audio is not enabled for this model")
        va = self._prepare(video, False) if video is not None else None
        aa = self._prepare(audio, True) if audio is not None else None
        return va, aa

    def _process_transformer_blocks(self, video, audio):
        for block in self.transformer_blocks.values():
            video, audio = block(video=video, audio=audio, inplace=True, workspace=self.workspace,
                                 seq_parallel=self.seq_parallel)
        return video, audio

    # ------------------------------------------------------------------ context cache / CUDA graphs (row N1)
    @staticmethod
    def _tensor_key(t: Tensor):
        return (t.data_ptr(), tuple(t.shape), tuple(t.stride()), t.dtype, t._version)

    def _select_cache(self, pre: str, key) -> Tuple[int, ContextCache, bool]:
        """The cache slot that holds (or will hold) the projections of the context named ``key``:
        -> (slot index, cache, hit).  A miss re-targets the least recently used slot (its buffers keep their
        addresses, only ``valid`` drops) — so ``cache.key`` always names what the buffers hold or are about to be
        filled with, on the eager path and under graph replay alike."""
        slots: List[ContextCache] = self._context_caches.setdefault(pre, [])
        for c in slots:
            if c.key == key:
                c.stamp = self._cache_clock = getattr(self, "_cache_clock", 0) + 1
                return slots.index(c), c, c.valid
        free = [c for c in slots if c.key is None]  # invalidated (a finished denoise loop): reuse before growing
        if free:
            c = free[0]
        elif len(slots) < self.context_cache_slots:
            slots.append(ContextCache())
            c = slots[-1]
        else:
            c = min(slots, key=lambda s_: s_.stamp)
        c.retarget(key)
        c.stamp = self._cache_clock = getattr(self, "_cache_clock", 0) + 1
        return slots.index(c), c, False

    def _context_cache_for(self, pre: str, context: Tensor) -> ContextCache:
        # under graph replay the forward sees a static copy of the context; the caller's tensor names the cache
        return self._select_cache(pre, self._context_keys.get(pre) or self._tensor_key(context))[1]

    def invalidate_context(self) -> None:
        """A new denoise loop starts: whatever prompt comes next is projected afresh (buffers and graphs are kept)."""
        for slots in self._context_caches.values():
            for c in slots:
                c.key, c.valid = None, False

    def clear_caches(self) -> None:
        self._context_caches.clear()
        if self._graphs is not None:
            self._graphs.clear()

    def __call__(self, video: Optional[Modality] = None, audio: Optional[Modality] = None):
        # sequence-parallel forwards are captured only when every exchange is a libltxb kernel on NVLink peer memory
        # (capturing NCCL collectives hung, profiles/README.md)
        sp_ok = self.seq_parallel is None or getattr(self.seq_parallel, "peers", None) is not None
        if self._graphs is not None and sp_ok and not torch.cuda.is_current_stream_capturing():
            return self._graphed_call(video, audio)
        return self._forward(video, audio)

    def _graphed_call(self, video: Optional[Modality], audio: Optional[Modality]):
        """One captured CUDA graph per input signature (and per context-cache state); the inputs are copied
        into the graph's static tensors, the outputs cloned out of them."""
        def fields(m):
            if m is None:
                return []
            return [m.latent, m.timesteps, m.positions, m.context, m.context_mask]

        def sig(t):
            return None if t is None else (tuple(t.shape), t.dtype)

        def pe_key(m):
            # precomputed RoPE tables are constant across a denoise loop and large (2 x 115 MB at 14 080 tokens): the graph
            # reads the CALLER's tensors in place (no static copy, no per-replay copy) and is keyed on their identity
            pe = None if m is None else m.positional_embeddings
            return None if pe is None else (pe[0].data_ptr(), pe[1].data_ptr(), tuple(pe[0].shape))

        flat = fields(video) + fields(audio)
        for t in flat:
            if t is not None and not t.is_cuda:
                raise LtxbError("LTXModel inputs must be CUDA tensors; there is no CPU fallback on this path")
        used = []  # (prefix, slot index, cache, hit) of the context caches this call goes through
        for pre, m in (("", video), ("audio_", audio)):
            if self.cache_context and m is not None:
                self._context_keys[pre] = self._tensor_key(m.context)
                slot, c, hit = self._select_cache(pre, self._context_keys[pre])
                used.append((pre, slot, c, hit))
        enabled = tuple(None if m is None else bool(m.enabled) for m in (video, audio))
        key = (video is None, audio is None, tuple(sig(t) for t in flat), tuple((pre, slot, hit) for pre, slot, _, hit in used),
               self.seq_parallel is not None, enabled, pe_key(video), pe_key(audio))
        try:
            entry = self._graphs.get(key)
            if entry is None:
                def static(m):
                    if m is None:
                        return None
                    c = lambda t: None if t is None else t.clone()  # noqa: E731
                    return Modality(c(m.latent), c(m.timesteps), c(m.positions), c(m.context), m.enabled, c(m.context_mask),
                                    m.positional_embeddings)
                sv, sa = static(video), static(audio)
                self._forward(sv, sa)  # eager warm-up: sizes the workspaces, configures the kernels
                torch.cuda.synchronize()
                for _, _, c, hit in used:  # capture the same variant (fill vs reuse) as this key names
                    c.valid = hit
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    out = self._forward(sv, sa)
                entry = self._graphs[key] = (graph, sv, sa, out)  # sv / sa also keep the callers' RoPE tables alive
            graph, sv, sa, out = entry
            for dst, src in zip(fields(sv) + fields(sa), flat):
                if dst is not None:
                    dst.copy_(src)
            graph.replay()
            for _, _, c, _ in used:  # c.key was set by _select_cache: key and contents move together
                c.valid = c.context is not None
        finally:
            self._context_keys.clear()
        return tuple(None if o is None else o.clone() for o in out)

    def _forward(self, video: Optional[Modality] = None, audio: Optional[Modality] = None):
        sp = self.seq_parallel
        if sp is not None:
            video, audio = sp.shard_inputs(video, audio)
        va, aa = self.prepare(video, audio)
        va, aa = self._process_transformer_blocks(va, aa)
        vdt = None if video is None else (BF16 if video.latent.dtype == BF16 else F32)
        adt = None if audio is None else (BF16 if audio.latent.dtype == BF16 else F32)
        vx = self._process_output(self.scale_shift_table, self.proj_out, va, vdt) if va is not None else None
        ax = self._process_output(self.audio_scale_shift_table, self.audio_proj_out, aa, adt) if aa is not None else None
        if sp is not None:
            vx, ax = sp.gather_outputs(vx, ax)
        for pre, m in (("", video), ("audio_", audio)):  # every block has now filled its K/V entry
            if self.cache_context and m is not None:
                c = self._context_cache_for(pre, m.context)
                c.valid = c.context is not None
        return vx, ax

    # ------------------------------------------------------------------ accounting
    def forward_flops(self, T: int, Tc: int, Ta: int = 0, B: int = 1) -> float:
        """Algorithmic FLOPs of one forward as the reference executes it (per-token AdaLN included)."""
        c = self.config
        L = c.num_layers
        f = 0.0
        if c.model_type.is_video_enabled():
            D = c.inner_dim
            f += L * (28.0 * T * D * D + 4.0 * T * T * D + 4.0 * Tc * D * D + 4.0 * T * Tc * D)
            f += 2.0 * T * (256 * D + D * D + 6 * D * D) + 2.0 * Tc * (c.caption_channels * D + D * D) + 2.0 * T * 2 * c.in_channels * D
        if c.model_type.is_audio_enabled():
            Da = c.audio_inner_dim
            f += L * (28.0 * Ta * Da * Da + 4.0 * Ta * Ta * Da + 4.0 * Tc * Da * Da + 4.0 * Ta * Tc * Da)
            f += 2.0 * Ta * (256 * Da + Da * Da + 6 * Da * Da) + 2.0 * Tc * (c.audio_caption_channels * Da + Da * Da) + 2.0 * Ta * 2 * c.audio_in_channels * Da
        if self._av:
            D, Da = c.inner_dim, c.audio_inner_dim
            a2v = 2.0 * T * D * Da + 4.0 * Ta * Da * Da + 4.0 * T * Ta * Da + 2.0 * T * Da * D
            v2a = 2.0 * Ta * Da * Da + 4.0 * T * D * Da + 4.0 * T * Ta * Da + 2.0 * Ta * Da * Da
            f += L * (a2v + v2a)
        return f * B


class X0Model:
    """ltx.py:888-906 — wraps a velocity model, returns denoised x0 = x - sigma * v (utils.py:404-440, fp32)."""

    def __init__(self, velocity_model: LTXModel) -> None:
        self.velocity_model = velocity_model

    def __call__(self, video: Optional[Modality] = None, audio: Optional[Modality] = None):
        vx, ax = self.velocity_model(video, audio)
        dv = to_denoised(video.latent, vx, video.timesteps) if vx is not None else None
        da = to_denoised(audio.latent, ax, audio.timesteps) if ax is not None else None
        return dv, da


def to_denoised(noisy: Tensor, velocity: Tensor, sigma) -> Tensor:
    """utils.py:404-440: x0 = x - sigma * v in fp32, result in ``noisy``'s dtype.  ``sigma`` is a float
    or a per-token tensor (B, T) / (B, 1)."""
    if not noisy.is_cuda:
        raise LtxbError("to_denoised needs CUDA tensors")
    B, T, C = noisy.shape
    x = noisy.to(F32).contiguous().clone()
    v = velocity.to(F32).contiguous()
    x0 = torch.empty_like(x)
    if isinstance(sigma, (int, float)):
        ops.euler_step(x.view(B * T, C), v.view(B * T, C), 1.0, 0.0, sigma_tok=torch.full((B * T,), float(sigma), dtype=F32, device=x.device), x0_out=x0.view(B * T, C))
    else:
        s = sigma.to(F32)
        s = s.reshape(B, -1).expand(B, T).contiguous().view(B * T)
        ops.euler_step(x.view(B * T, C), v.view(B * T, C), 1.0, 0.0, sigma_tok=s, x0_out=x0.view(B * T, C))
    return x0.to(noisy.dtype)
