"""ORACLE / test infrastructure only — builds the small MLX-quantised LTX-2 checkpoints that pin row N3
(quantised half) and row N2 (checkpoint ingest): seeded weights of ``small_config`` are quantised with the oracle's
restatement of ``mx.quantize`` and written as safetensors files in the two on-disk layouts the reference accepts
(ltx.py:548-564, 614-632):

* ``mlx``      — already-sanitised names as ``convert.py`` saves them (no prefix), 4 bits / group 64, every linear of
                 the transformer blocks quantised ("core" scope, convert.py:728-739), NO quantization.json, so the
                 reference falls back to its defaults group_size=64 / bits=4 (ltx.py:645-648);
* ``upstream`` — ``model.diffusion_model.`` prefix with ``.to_out.0.`` / ``.ff.net.0.proj.`` / ``.ff.net.2.`` /
                 ``.linear_1.`` names, 8 bits / group 32, only ``attn1`` quantised, a quantization.json beside it
                 (ltx.py:649-668), F32 tables and biases (the loader rounds them to bf16, ltx.py:613-615), and extra
                 tensors a video-only model must ignore (audio twins, VAE, embeddings connectors; ltx.py:739-740).

Everything is regenerated from the seed on both sides (generator and tests), so only the reference's outputs are
committed (tests/golden/quant.npz).  Used by oracle/make_golden_quant.py and tests/ — never by the product.
"""
from __future__ import annotations

import json
import struct
import zlib
from pathlib import Path
from typing import Dict, Tuple

import numpy as np
import torch

import ltx_oracle as O

VARIANTS = {
    "mlx": dict(bits=4, group_size=64, scope="core", prefixed=False, meta=False, seed=11),
    "upstream": dict(bits=8, group_size=32, scope="attn1", prefixed=True, meta=True, seed=12),
}
GRID, TC, SIGMA = (2, 8, 8), 48, 0.725  # 128 video tokens, 48 text tokens
# LoRA applied on top of the quantised model (the reference attaches runtime adapters there, lora.py:188-275 via
# generate.py:2999-3023): on-disk key prefix -> (sanitised weight name, out, in) at small_config widths
LORA_TARGETS = [
    ("diffusion_model.transformer_blocks.0.attn1.to_q", "transformer_blocks.0.attn1.to_q.weight", 512, 512),
    ("diffusion_model.transformer_blocks.1.attn1.to_out.0", "transformer_blocks.1.attn1.to_out.weight", 512, 512),
    ("diffusion_model.transformer_blocks.1.ff.net.0.proj", "transformer_blocks.1.ff.proj_in.weight", 2048, 512),
    ("diffusion_model.transformer_blocks.0.attn2.to_k", "transformer_blocks.0.attn2.to_k.weight", 512, 512),
]
LORA_RANK, LORA_STRENGTH = 8, 0.7


def config() -> O.OracleConfig:
    return O.small_config(O.LTXModelType.VideoOnly, num_layers=2)


def _in_scope(name: str, scope: str) -> bool:
    """convert.py:716-739: which linears the converter quantises."""
    if "transformer_blocks" not in name:
        return False
    if scope == "attn1":
        return ".attn1." in name
    return ".attn" in name or ".ff." in name


def build(variant: str) -> Tuple[Dict[str, object], Dict[str, torch.Tensor]]:
    """-> (state dict with quantised linears: uint32 ndarrays + bf16 scales/biases, the same weights dequantised to
    fp32 for OracleLTXModel).  Non-quantised tensors carry the values the reference holds AFTER its load-time cast
    (fp32 -> bf16, ltx.py:613-615), kept as fp32 here."""
    v = VARIANTS[variant]
    cfg = config()
    tensors = O.init_params(cfg, seed=v["seed"])
    state: Dict[str, object] = {}
    for k, t in tensors.items():
        t16 = t.to(torch.bfloat16)
        if k.endswith(".weight") and t.dim() == 2 and _in_scope(k, v["scope"]):
            packed, s, b = O.affine_quantize(t16, v["group_size"], v["bits"])
            base = k[: -len(".weight")]
            state[k], state[base + ".scales"], state[base + ".biases"] = packed, s, b
        else:
            state[k] = t16.float()
    in_features = {k: t.shape[1] for k, t in tensors.items() if t.dim() == 2}
    return state, O.dequantize_state_dict(state, in_features)


def _upstream_name(k: str) -> str:
    for new, old in ((".to_out.0.", ".to_out."), (".ff.net.0.proj.", ".ff.proj_in."), (".ff.net.2.", ".ff.proj_out."),
                     (".linear_1.", ".linear1."), (".linear_2.", ".linear2.")):
        k = k.replace(old, new)
    return "model.diffusion_model." + k


def write_safetensors(path: Path, entries: Dict[str, Tuple[str, tuple, bytes]]) -> None:
    """Minimal safetensors writer: 8-byte little-endian header length, JSON header, raw little-endian data."""
    header, blobs, off = {}, [], 0
    for name, (dtype, shape, raw) in entries.items():
        header[name] = {"dtype": dtype, "shape": list(shape), "data_offsets": [off, off + len(raw)]}
        blobs.append(raw)
        off += len(raw)
    hj = json.dumps(header).encode()
    hj += b" " * (-len(hj) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hj)))
        f.write(hj)
        for b in blobs:
            f.write(b)


def _entry(value, f32: bool = False) -> Tuple[str, tuple, bytes]:
    if isinstance(value, np.ndarray):  # packed levels
        return "U32", value.shape, value.astype("<u4").tobytes()
    t = value.detach().contiguous()
    if f32:
        return "F32", tuple(t.shape), t.float().numpy().astype("<f4").tobytes()
    return "BF16", tuple(t.shape), t.to(torch.bfloat16).view(torch.int16).numpy().astype("<i2").tobytes()


def write_checkpoint(variant: str, directory: Path) -> Path:
    """Writes ``<directory>/<variant>.safetensors`` (+ quantization.json for the upstream variant)."""
    v = VARIANTS[variant]
    state, _ = build(variant)
    g = torch.Generator().manual_seed(99)
    entries: Dict[str, Tuple[str, tuple, bytes]] = {}
    for k, val in state.items():
        name = _upstream_name(k) if v["prefixed"] else k
        # upstream variant: 1-D tensors and tables on disk as F32 (already bf16-exact values, so the loader's cast is exact)
        f32 = v["prefixed"] and isinstance(val, torch.Tensor) and (val.dim() == 1 or "scale_shift_table" in k) \
            and not (k.endswith(".scales") or k.endswith(".biases"))
        entries[name] = _entry(val, f32)
    if v["prefixed"]:  # tensors the video-only transformer must ignore
        entries["model.diffusion_model.audio_patchify_proj.weight"] = _entry(torch.randn(64, 128, generator=g))
        entries["model.diffusion_model.video_embeddings_connector.proj.weight"] = _entry(torch.randn(8, 8, generator=g))
        entries["vae.decoder.conv_in.weight"] = _entry(torch.randn(4, 4, generator=g))
    else:
        entries["audio_patchify_proj.weight"] = _entry(torch.randn(64, 128, generator=g))
    directory.mkdir(parents=True, exist_ok=True)
    path = directory / f"{variant}.safetensors"
    write_safetensors(path, entries)
    meta = directory / "quantization.json"
    if v["meta"]:
        meta.write_text(json.dumps({"group_size": v["group_size"], "bits": v["bits"], "mode": "affine",
                                    "predicate": "attn1_only", "dtype": "bfloat16"}))
    elif meta.exists():
        meta.unlink()
    return path


def inputs(variant: str) -> O.Modality:
    g = torch.Generator().manual_seed(VARIANTS[variant]["seed"] + 100)
    cfg = config()
    F_, H_, W_ = GRID
    T = F_ * H_ * W_
    lat = torch.randn(1, T, cfg.in_channels, generator=g)
    ctx = torch.randn(1, TC, cfg.caption_channels, generator=g)
    pos = torch.from_numpy(O.create_position_grid(1, F_, H_, W_))
    return O.Modality(lat, torch.full((1, T), SIGMA), pos, ctx, True, None)


def packed_checksum(state: Dict[str, object]) -> int:
    """crc32 over every packed tensor + scales + biases, in name order: detects any drift of the regenerated checkpoint."""
    crc = 0
    for k in sorted(state):
        val = state[k]
        if isinstance(val, np.ndarray):
            crc = zlib.crc32(val.astype("<u4").tobytes(), crc)
        elif k.endswith(".scales") or k.endswith(".biases"):
            crc = zlib.crc32(val.to(torch.bfloat16).view(torch.int16).numpy().tobytes(), crc)
    return crc


def lora_state(variant: str) -> Dict[str, torch.Tensor]:
    """Seeded LoRA A / B pairs (bf16-exact values held in fp32) under their on-disk key names."""
    g = torch.Generator().manual_seed(VARIANTS[variant]["seed"] + 200)
    sd: Dict[str, torch.Tensor] = {}
    for prefix, _, out_f, in_f in LORA_TARGETS:
        sd[f"{prefix}.lora_A.weight"] = (0.15 * torch.randn(LORA_RANK, in_f, generator=g)).to(torch.bfloat16).float()
        sd[f"{prefix}.lora_B.weight"] = (0.15 * torch.randn(out_f, LORA_RANK, generator=g)).to(torch.bfloat16).float()
    return sd


def write_lora(variant: str, directory: Path) -> Path:
    path = directory / f"{variant}_lora.safetensors"
    write_safetensors(path, {k: _entry(v, f32=True) for k, v in lora_state(variant).items()})
    return path
