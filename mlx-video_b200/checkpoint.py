"""Checkpoint ingest: safetensors file(s) -> the state-dict names ``LTXModel.load_weights``
expects.  Mirrors the key mapping of the reference's ``LTXModel.sanitize`` / ``from_pretrained``
(mlx_video/models/ltx/ltx.py:508-533, 548-564): only ``model.diffusion_model.*`` tensors belong to the
transformer; the embeddings connectors live in the text encoder; upstream names carry ``.to_out.0.``,
``.ff.net.0.proj.``, ``.ff.net.2.``, ``.linear_1.`` / ``.linear_2.``.  Already-sanitised files (no prefix) load
as they are.  Pre-quantised MLX checkpoints (``.scales`` / ``.biases`` siblings, ltx.py:641-725) are rejected:
quantised linears (row N3) are not built; LoRA files are handled by ``lora.py``.
"""
from __future__ import annotations

import json
import struct
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Union

import torch

PREFIX = "model.diffusion_model."
_RENAMES = [(".to_out.0.", ".to_out."), (".ff.net.0.proj.", ".ff.proj_in."), (".ff.net.2.", ".ff.proj_out."),
            (".audio_ff.net.0.proj.", ".audio_ff.proj_in."), (".audio_ff.net.2.", ".audio_ff.proj_out."),
            (".linear_1.", ".linear1."), (".linear_2.", ".linear2.")]


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


def scan_keys(paths: Iterable[Path]) -> List[str]:
    """Tensor names from the safetensors headers, without touching tensor data (ltx.py:566-590)."""
    keys: List[str] = []
    for p in paths:
        with open(p, "rb") as f:
            (n,) = struct.unpack("<Q", f.read(8))
            header = json.loads(f.read(n))
        keys += [k for k in header if k != "__metadata__"]
    return keys


def load_transformer_weights(model_path: Union[str, Path, List[Path]], config=None) -> Dict[str, torch.Tensor]:
    """Read every transformer tensor of the file(s) (a directory means all ``*.safetensors`` in it)."""
    from safetensors import safe_open

    paths = model_path if isinstance(model_path, (list, tuple)) else [model_path]
    files: List[Path] = []
    for p in map(Path, paths):
        files += sorted(p.glob("*.safetensors")) if p.is_dir() else [p]
    if not files:
        raise FileNotFoundError(f"no safetensors files under {model_path}")
    names = scan_keys(files)
    if any(k.endswith(".scales") for k in names):
        raise ValueError("pre-quantised MLX checkpoint (.scales/.biases tensors): quantised linears are not supported on this path")
    out: Dict[str, torch.Tensor] = {}
    for f in files:
        with safe_open(str(f), framework="pt", device="cpu") as sf:
            for k in sf.keys():
                name = sanitize_key(k)
                if name is not None:
                    out[name] = sf.get_tensor(k)
    return out
