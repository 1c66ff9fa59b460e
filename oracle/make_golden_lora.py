"""ORACLE / test infrastructure only — generates tests/golden/lora.npz by running the reference's OWN
``mlx_video/lora.py::apply_lora_to_weights`` (lora.py:93-129; the path generate.py:3007 takes for non-quantised base
weights) unmodified over oracle/mlx_shim, on seeded weights and two LoRA files written with safetensors.

    python oracle/make_golden_lora.py        # needs /root/reference (this container); the fixture is committed

Inputs and outputs are stored as float32 images of bf16 tensors (npz has no bf16): base weights, the LoRA A / B
pairs under their on-disk key names, strengths, and the merged weights the reference returns.
"""
from __future__ import annotations

import importlib
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import ref_loader  # noqa: E402

GOLDEN = HERE.parent / "tests" / "golden"

# (on-disk LoRA key prefix, sanitised model weight name, out, in) — upstream prefixes / renames of lora.py:18-33
TARGETS = [
    ("diffusion_model.transformer_blocks.0.attn1.to_q", "transformer_blocks.0.attn1.to_q.weight", 256, 256),
    ("diffusion_model.transformer_blocks.0.attn1.to_out.0", "transformer_blocks.0.attn1.to_out.weight", 256, 256),
    ("model.diffusion_model.transformer_blocks.1.ff.net.0.proj", "transformer_blocks.1.ff.proj_in.weight", 1024, 256),
    ("transformer_blocks.1.ff.net.2", "transformer_blocks.1.ff.proj_out.weight", 256, 1024),
    ("diffusion_model.transformer_blocks.1.attn2.to_k", "transformer_blocks.1.attn2.to_k.weight", 256, 256),
    ("diffusion_model.not_in_model.proj", None, 64, 64),  # skipped by the reference: no such weight
]
RANKS, STRENGTHS = (16, 40), (0.8, -0.35)  # second rank is not a multiple of 16 / 64 on purpose


def main() -> int:
    ref_loader.load()  # registers the mlx shim and the package stubs
    lora = importlib.import_module("mlx_video.lora")
    mx = importlib.import_module("mlx.core")
    from safetensors.torch import save_file

    g = torch.Generator().manual_seed(2024)
    base = {name: (torch.randn(o, i, generator=g) / i ** 0.5).to(torch.bfloat16) for _, name, o, i in TARGETS if name}
    base["transformer_blocks.0.attn1.to_q.bias"] = torch.randn(256, generator=g)  # untouched by LoRA
    out = {"strengths": np.asarray(STRENGTHS, np.float32)}
    specs = []
    with tempfile.TemporaryDirectory() as td:
        for n, (rank, strength) in enumerate(zip(RANKS, STRENGTHS)):
            sd = {}
            for prefix, name, o, i in TARGETS:
                if n == 1 and name and "ff" in name:
                    continue  # the second LoRA only touches the attention projections
                sd[f"{prefix}.lora_A.weight"] = (0.3 * torch.randn(rank, i, generator=g)).to(torch.bfloat16)
                sd[f"{prefix}.lora_B.weight"] = (0.3 * torch.randn(o, rank, generator=g)).to(torch.bfloat16)
            path = Path(td) / f"lora{n}.safetensors"
            save_file({k: v.float() for k, v in sd.items()}, str(path))  # f32 on disk: bf16-exact values, numpy-readable
            for k, v in sd.items():
                out[f"lora{n}/{k}"] = v.float().numpy()
            specs.append(lora.LoraSpec(path, strength))
        merged = lora.apply_lora_to_weights({k: mx.array(v) for k, v in base.items()}, specs, verbose=True)
    for k, v in base.items():
        out[f"base/{k}"] = v.float().numpy()
    changed = 0
    for k, v in merged.items():
        assert v.dtype == base[k].dtype, (k, v.dtype)
        out[f"merged/{k}"] = v._t.float().numpy()
        changed += int(not torch.equal(v._t, base[k]))
    assert changed == 5, changed
    GOLDEN.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(GOLDEN / "lora.npz", **out)
    print(f"wrote {GOLDEN / 'lora.npz'}: {len(out)} arrays, {changed} weights changed")
    return 0


if __name__ == "__main__":
    sys.exit(main())
