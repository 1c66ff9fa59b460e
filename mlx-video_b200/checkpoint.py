"""Checkpoint ingest: safetensors file(s) -> the state-dict names ``LTXModel.load_weights``
expects.  Mirrors the reference's ``LTXModel.sanitize`` / ``from_pretrained`` (mlx_video/models/ltx/ltx.py:508-533,
535-885):

* key mapping (ltx.py:548-564): when any tensor name starts with ``model.diffusion_model.`` the file is an upstream
  (PyTorch-layout) checkpoint — only tensors with that prefix belong to the transformer, the embeddings connectors live
  in the text encoder, and names carry ``.to_out.0.``, ``.ff.net.0.proj.``, ``.ff.net.2.``, ``.linear_1.`` /
  ``.linear_2.``; otherwise names are taken as they are (files written by the reference's converter);
* header scan without touching tensor data (ltx.py:566-590);
* pre-quantised MLX checkpoints (ltx.py:614-725): a linear is quantised iff the file holds ``<linear>.scales``
  (the reference's ``_scales_predicate``); ``quantization.json`` beside the first file gives group_size / bits / the
  dtype of scales and biases (defaults 64 / 4, ltx.py:645-668).  Here the layout is read off the tensor shapes and
  the JSON, when present, has to agree.  The packed tensors are returned as they are; ``LTXModel.load_weights``
  expands them to bf16 on the device (``ltxb_dequant_affine_bf16``);
* the load-time cast (ltx.py:592-615): fp32 tensors are rounded to bf16 VALUES, except scales / biases.

LoRA files are handled by ``lora.py``.
"""
from __future__ import annotations

import json
import struct
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Set, Tuple, Union

import torch

PREFIX = "model.diffusion_model."
_RENAMES = [(".to_out.0.", ".to_out."), (".ff.net.0.proj.", ".ff.proj_in."), (".ff.net.2.", ".ff.proj_out."),
            (".audio_ff.net.0.proj.", ".audio_ff.proj_in."), (".audio_ff.net.2.", ".audio_ff.proj_out."),
            (".linear_1.", ".linear1."), (".linear_2.", ".linear2.")]
_AUX_DTYPES = {"bf16": torch.bfloat16, "bfloat16": torch.bfloat16, "f16": torch.float16, "float16": torch.float16,
               "fp16": torch.float16, "f32": torch.float32, "float32": torch.float32, "fp32": torch.float32}


def sanitize_key(key: str) -> Optional[str]:
    """Upstream checkpoint key -> model parameter name, or None for tensors that are not the transformer's."""
    if "audio_embeddings_connector" in key or "video_embeddings_connector" in key:
        return None
    if key.startswith(PREFIX):
        key = key[len(PREFIX):]
    elif "." in key and key.split(".")[0] in ("vae", "audio_vae", "vocoder", "text_encoder", "model"):
        return None
    for old, new in _RENAMES:
        key = key.replace(old, new)
    return key


def _is_aux(key: str) -> bool:
    return key.endswith(".scales") or key.endswith(".biases")


def _maybe_cast(key: str, value: torch.Tensor, aux_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """ltx.py:592-615: quantisation scales / biases keep their dtype (or take quantization.json's); every other fp32
    tensor is rounded to bf16."""
    if _is_aux(key):
        return value if aux_dtype is None or value.dtype == aux_dtype else value.to(aux_dtype)
    return value.to(torch.bfloat16) if value.dtype == torch.float32 else value


def sanitize_state_dict(weights: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """In-memory tensors (``weights_override``, ltx.py:617-623, 828-842): upstream names are mapped when present,
    then the load-time cast is applied."""
    if any(k.startswith(PREFIX) for k in weights):
        mapped = {}
        for k, v in weights.items():
            name = sanitize_key(k) if k.startswith(PREFIX) else None
            if name is not None:
                mapped[name] = v
        weights = mapped
    return {k: _maybe_cast(k, v) for k, v in weights.items()}


def scan_keys(paths: Iterable[Path]) -> List[str]:
    """Tensor names from the safetensors headers, without touching tensor data (ltx.py:566-590)."""
    keys: List[str] = []
    for p in paths:
        with open(p, "rb") as f:
            (n,) = struct.unpack("<Q", f.read(8))
            header = json.loads(f.read(n))
        keys += [k for k in header if k != "__metadata__"]
    return keys


def read_quantization_meta(first_file: Path) -> Dict[str, object]:
    """``quantization.json`` beside the first weight file (ltx.py:649-668); {} when absent or unreadable."""
    meta_path = Path(first_file).parent / "quantization.json"
    try:
        if meta_path.exists():
            with open(meta_path, "r") as f:
                meta = json.load(f)
            return meta if isinstance(meta, dict) else {}
    except Exception:  # the reference ignores a broken file as well
        pass
    return {}


def quant_layout(name: str, packed_shape: Tuple[int, ...], scales_shape: Tuple[int, ...], dense_shape: Tuple[int, ...],
                 meta: Optional[Dict[str, object]] = None) -> Tuple[int, int]:
    """(group_size, bits) of one MLX affine-quantised linear, from its tensor shapes: packed (out, in*bits/32) uint32,
    scales (out, in/group_size).  ValueError when the shapes do not describe ``dense_shape`` or contradict ``meta``."""
    out_f, in_f = dense_shape
    ok = len(packed_shape) == 2 and len(scales_shape) == 2 and packed_shape[0] == out_f == scales_shape[0] \
        and packed_shape[1] > 0 and scales_shape[1] > 0 and (packed_shape[1] * 32) % in_f == 0 and in_f % scales_shape[1] == 0
    if not ok:
        raise ValueError(f"shape mismatch for quantised {name}: packed {packed_shape}, scales {scales_shape} vs model {dense_shape}")
    bits, group_size = packed_shape[1] * 32 // in_f, in_f // scales_shape[1]
    if bits not in (2, 4, 8) or group_size not in (32, 64, 128):
        raise ValueError(f"unsupported quantisation of {name}: {bits} bits, group size {group_size} (2/4/8 bits, groups of 32/64/128)")
    if meta:
        if str(meta.get("mode", "affine")) != "affine":
            raise ValueError(f"unsupported quantisation mode {meta.get('mode')!r} (affine only)")
        want = (int(meta.get("group_size", group_size)), int(meta.get("bits", bits)))
        if want != (group_size, bits):
            raise ValueError(f"quantization.json says group_size/bits {want} but {name} is stored with {(group_size, bits)}")
    return group_size, bits


def checkpoint_files(model_path: Union[str, Path, List[Path]]) -> List[Path]:
    """A file, a list of files, or a directory (all ``*.safetensors`` in it)."""
    paths = model_path if isinstance(model_path, (list, tuple)) else [model_path]
    files: List[Path] = []
    for p in map(Path, paths):
        files += sorted(p.glob("*.safetensors")) if p.is_dir() else [p]
    if not files:
        raise FileNotFoundError(f"no safetensors files under {model_path}")
    return files


def load_transformer_weights(model_path: Union[str, Path, List[Path]], config=None,
                             expected: Optional[Set[str]] = None) -> Dict[str, torch.Tensor]:
    """Read the transformer tensors of the file(s) (a directory means all ``*.safetensors`` in it), under model
    parameter names.  ``expected``: keep only these names (the reference's ``_should_load_key``, ltx.py:739-740).
    Packed uint32 weights and their scales / biases come back as stored."""
    from safetensors import safe_open

    files = checkpoint_files(model_path)
    names = scan_keys(files)
    is_upstream = any(k.startswith(PREFIX) for k in names)

    def mapped(k: str) -> Optional[str]:
        if is_upstream:  # ltx.py:548-564: anything without the prefix is not the transformer's
            return sanitize_key(k) if k.startswith(PREFIX) else None
        return sanitize_key(k)

    has_quant = any(_is_aux(m) for m in map(mapped, names) if m)
    meta = read_quantization_meta(files[0]) if has_quant else {}
    aux_dtype = _AUX_DTYPES.get(str(meta.get("dtype", "")).lower().strip())
    out: Dict[str, torch.Tensor] = {}
    for f in files:
        with safe_open(str(f), framework="pt", device="cpu") as sf:
            for k in sf.keys():
                name = mapped(k)
                if name is None or (expected is not None and name not in expected):
                    continue
                out[name] = _maybe_cast(name, sf.get_tensor(k), aux_dtype)
    return out
