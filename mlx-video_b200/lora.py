"""LoRA on the B200 path.

Mirror of the reference's ``mlx_video/lora.py``: same names and argument meaning — ``LoraSpec``, ``load_lora_state``,
``has_quantized_weights``, ``apply_lora_to_weights`` (dict in, dict out) and ``apply_lora_to_model`` (an already loaded
model, in place).  Both of the reference's modes end in the same thing here, a merged bf16 device weight:

* non-quantised checkpoints — the reference merges every LoRA into the weights before the model is built
  (``apply_lora_to_weights``, lora.py:93-129, called at generate.py:3007): w = bf16(w + bf16((B.A) * strength));
* quantised checkpoints — the reference must not touch the packed weights, so it wraps the linears in runtime
  ``LoRAAdapter``s, y + ((x A^T) B^T) * strength (lora.py:188-275, generate.py:2999-3023).  On this path quantised
  linears were already expanded to bf16 at load (``ltxb_dequant_affine_bf16``, checkpoint.py), nothing is re-quantised,
  and ``apply_lora_to_model`` merges into the expanded weight: the same sum W x + s B A x with one bf16 rounding of
  W + s B A instead of one of the adapter's output; it costs nothing per denoise step (pinned by
  tests/golden/quant.npz ``velocity_lora``, produced by the reference's adapters).

A merged weight is the right trade on a GPU that holds the 25.8 GB of weights resident: the stage-2 distilled LoRA of
the two-stage pipelines is merged once at stage switch.  The arithmetic runs through the C ABI: delta = B . A on the
tensor cores (``ltxb_gemm_bf16``, fp32 accumulation and output), then ``ltxb_lora_merge_bf16`` applies the reference's
two roundings.  A and B are taken in bf16 (LTX-2 LoRA files are bf16; fp32 files lose their extra mantissa bits here).
"""
from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path
from typing import Dict, Iterable, Iterator, List, Optional, Tuple

import torch

from . import _lib, ops

Tensor = torch.Tensor
_RENAMES = ((".to_out.0.", ".to_out."), (".ff.net.0.proj.", ".ff.proj_in."), (".ff.net.2.", ".ff.proj_out."),
            (".audio_ff.net.0.proj.", ".audio_ff.proj_in."), (".audio_ff.net.2.", ".audio_ff.proj_out."),
            (".linear_1.", ".linear1."), (".linear_2.", ".linear2."))


@dataclass(frozen=True)
class LoraSpec:
    """lora.py:12-15."""

    path: Path
    strength: float = 1.0


def _sanitize_lora_prefix(prefix: str) -> str:
    """lora.py:18-33."""
    for p in ("model.diffusion_model.", "diffusion_model."):
        if prefix.startswith(p):
            prefix = prefix[len(p):]
    for old, new in _RENAMES:
        prefix = prefix.replace(old, new)
    return prefix


def _candidate_weight_keys(base_raw: str, base_sanitized: str) -> Tuple[str, ...]:
    """lora.py:72-90."""
    cand = [base_sanitized, base_raw]
    if base_raw.startswith("diffusion_model."):
        cand.append(f"model.{base_raw}")
    if base_sanitized and not base_sanitized.startswith("model."):
        cand += [f"diffusion_model.{base_sanitized}", f"model.diffusion_model.{base_sanitized}"]
    return tuple(dict.fromkeys(cand))


def load_lora_state(path: Path) -> Dict[str, Tensor]:
    """lora.py:48-53: every tensor of a LoRA safetensors file (CPU)."""
    from safetensors import safe_open

    out: Dict[str, Tensor] = {}
    with safe_open(str(path), framework="pt", device="cpu") as f:
        for k in f.keys():
            out[k] = f.get_tensor(k)
    return out


def _iter_lora_pairs(lora_sd: Dict[str, Tensor]) -> Iterator[Tuple[str, str, Tensor, Tensor]]:
    """lora.py:56-69: (base key raw, base key sanitised, A (r, in), B (out, r))."""
    for key in lora_sd:
        if not key.endswith(".lora_A.weight"):
            continue
        key_b = key[: -len(".lora_A.weight")] + ".lora_B.weight"
        if key_b not in lora_sd:
            continue
        base = key.replace(".lora_A.weight", ".weight")
        yield base, _sanitize_lora_prefix(base), lora_sd[key], lora_sd[key_b]


def has_quantized_weights(weights: Dict[str, Tensor]) -> bool:
    """lora.py:132-133."""
    return any(k.endswith(".scales") or k.endswith(".biases") for k in weights)


def merge_lora_pair(weight: Tensor, A: Tensor, B: Tensor, strength: float) -> None:
    """weight (bf16 CUDA, (out, in), any row stride) <- bf16(weight + bf16((B @ A) * strength)), in place."""
    if not weight.is_cuda or weight.dtype != torch.bfloat16:
        raise _lib.LtxbError("merge_lora_pair needs a bf16 CUDA weight; there is no CPU fallback on this path")
    out_f, in_f = weight.shape
    r = A.shape[0]
    if A.dim() != 2 or B.dim() != 2 or A.shape[1] != in_f or B.shape != (out_f, r):
        raise ValueError(f"LoRA shapes A {tuple(A.shape)} / B {tuple(B.shape)} do not fit a weight {tuple(weight.shape)}")
    dev = weight.device
    rp = (r + 63) // 64 * 64  # the GEMM contracts over K in blocks of 64: zero-pad the rank (layout plumbing, load time)
    b_op = torch.zeros(out_f, rp, dtype=torch.bfloat16, device=dev)
    b_op[:, :r] = B.to(dev, torch.bfloat16)
    a_op = torch.zeros(in_f, rp, dtype=torch.bfloat16, device=dev)  # A^T: the GEMM's W operand is (N, K)
    a_op[:, :r] = A.to(dev, torch.bfloat16).t()
    delta = torch.empty(out_f, in_f, dtype=torch.float32, device=dev)
    ops.gemm(b_op, a_op, None, delta, mode=_lib.EPI_BIAS_F32)
    ops.lora_merge(weight, delta, float(strength))


def apply_lora_to_weights(weights: Dict[str, Tensor], lora_specs: Iterable[LoraSpec], verbose: bool = False,
                          device="cuda") -> Dict[str, Tensor]:
    """lora.py:93-129: a new dict with every LoRA of ``lora_specs`` merged, in order, into the weights they name.
    Touched weights come back as bf16 tensors on ``device`` (what ``LTXModel.load_weights`` stores anyway)."""
    if has_quantized_weights(weights):  # generate.py:2999-3004 never merges into packed weights
        raise ValueError("quantised weights (.scales / .biases): load the model first, then apply_lora_to_model()")
    updated = dict(weights)
    for spec in lora_specs:
        lora_sd = load_lora_state(spec.path)
        applied = skipped = 0
        for base_raw, base_san, A, B in _iter_lora_pairs(lora_sd):
            key = next((k for k in _candidate_weight_keys(base_raw, base_san) if k in updated), None)
            if key is None or updated[key].dim() != 2:
                skipped += 1
                continue
            w = updated[key]
            if not (w.is_cuda and w.dtype == torch.bfloat16 and w is not weights.get(key)):
                w = w.to(device=device, dtype=torch.bfloat16, copy=True)  # never write into the caller's tensors
            merge_lora_pair(w, A, B, spec.strength)
            updated[key] = w
            applied += 1
        _report(spec, applied, skipped, verbose)
    return updated


def apply_lora_to_model(model, lora_specs: Iterable[LoraSpec], verbose: bool = False):
    """Merge every LoRA into the loaded model's device weights IN PLACE (same key matching as
    ``apply_lora_to_weights``; fused q|k|v and stacked text-K/V storage are row views, so they are updated too).
    Cross-step caches of projected context are invalidated."""
    params = model.parameters()
    total = 0
    for spec in lora_specs:
        lora_sd = load_lora_state(spec.path)
        applied = skipped = 0
        for base_raw, base_san, A, B in _iter_lora_pairs(lora_sd):
            key = next((k for k in _candidate_weight_keys(base_raw, base_san) if k in params), None)
            if key is None or params[key].dim() != 2 or params[key].dtype != torch.bfloat16:
                skipped += 1
                continue
            merge_lora_pair(params[key], A, B, spec.strength)
            applied += 1
        total += applied
        _report(spec, applied, skipped, verbose)
    if total:
        model.clear_caches()  # projected-context caches (and the graphs that fill them) were computed with the old weights
        from .packed import REGISTRY as packed_registry

        packed_registry.clear()  # the merged weights exist in bf16 only: packed copies kept by load_weights(keep_packed=True) are stale
    return model


def _report(spec: LoraSpec, applied: int, skipped: int, verbose: bool) -> None:
    if verbose:
        print(f"[LoRA] {spec.path} applied={applied} skipped={skipped}")
    elif applied == 0:
        print(f"[LoRA] Warning: no weights applied for {spec.path}. Check key mapping.")


__all__: List[str] = ["LoraSpec", "load_lora_state", "apply_lora_to_weights", "apply_lora_to_model", "merge_lora_pair",
                      "has_quantized_weights"]
