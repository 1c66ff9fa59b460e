"""ORACLE / test infrastructure only — seeded cases of the latent-upsampler parity tests (tests/golden/upsampler.npz).
Shared by oracle/make_golden_upsampler.py (runs the reference's own upsampler on them) and tests/."""
from __future__ import annotations

from typing import Dict

import torch

import upsampler_oracle as U

# case -> widths, seeds, latent shape (B, 128, F, H, W)
CASES: Dict[str, dict] = {
    "small": dict(mid=128, blocks=2, seed=41, shape=(1, 128, 2, 4, 6)),   # non-square grid
    "b2_f1": dict(mid=128, blocks=1, seed=42, shape=(2, 128, 1, 3, 5)),  # batch 2, a single frame, 4 channels per group
    "loaded": dict(mid=128, blocks=4, seed=43, shape=(1, 128, 3, 4, 4)),  # through load_upsampler (always 4 blocks per stage)
}


def params(case: str) -> Dict[str, torch.Tensor]:
    c = CASES[case]
    p = U.init_upsampler_params(128, c["mid"], c["blocks"], seed=c["seed"])
    # conv weights rounded to bf16 values: the GPU path stores them in bf16, the comparison measures arithmetic only
    return {k: (v.to(torch.bfloat16).float() if k.endswith("weight") and v.dim() >= 4 else v) for k, v in p.items()}


def inputs(case: str):
    """-> (latent, latent_mean, latent_std): normalised latents and the VAE's per-channel statistics."""
    c = CASES[case]
    g = torch.Generator().manual_seed(c["seed"] + 500)
    latent = torch.randn(*c["shape"], generator=g)
    mean = 0.3 * torch.randn(128, generator=g)
    std = 0.5 + torch.rand(128, generator=g)
    return latent, mean, std


def upstream_state(case: str) -> Dict[str, torch.Tensor]:
    """The same weights in the upstream (PyTorch) layouts load_upsampler expects: conv3d (O, I, D, H, W), conv2d (O, I, H, W)."""
    out = {}
    for k, v in params(case).items():
        if v.dim() == 5:
            v = v.permute(0, 4, 1, 2, 3)
        elif v.dim() == 4:
            v = v.permute(0, 3, 1, 2)
        out[k] = v.contiguous()
    return out


def run_oracle(case: str) -> torch.Tensor:
    latent, mean, std = inputs(case)
    return U.upsample_latents(latent, params(case), mean, std)
