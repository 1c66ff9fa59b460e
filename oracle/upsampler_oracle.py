"""ORACLE — test infrastructure, not product code.

CPU restatement (PyTorch fp32) of the reference's latent 2x spatial upsampler, the step between the two stages of the
distilled / dev pipelines (first half): ``mlx_video/models/ltx/upsampler.py`` — ``Conv3d`` :6-72,
``GroupNorm3d`` :75-114, ``PixelShuffle2D`` :117-139, ``SpatialRationalResampler`` :142-174, ``ResBlock3D`` :177-199,
``LatentUpsampler`` :202-294, ``upsample_latents`` :297-316, ``load_upsampler`` :319-373.  Pinned by
tests/golden/upsampler.npz, which oracle/make_golden_upsampler.py produced by running that file itself over the shim.
What is restated rather than executed is the MLX library arithmetic: ``mx.conv3d`` / ``nn.Conv2d`` (channels-last
cross-correlation, weights (out, *kernel, in)), ``mx.mean`` / ``mx.var`` (population variance), ``nn.silu``.

Parameters use the reference's own names and MLX layouts: ``initial_conv.weight`` (O, 3, 3, 3, I), ``*.norm*.weight``,
``res_blocks.{i}.conv{1,2}.*``, ``upsampler.conv.weight`` (4*mid, 3, 3, mid), ``post_upsample_res_blocks.{i}.*``,
``final_conv.*``.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
NUM_GROUPS, GN_EPS = 32, 1e-5  # upsampler.py:77,182,184,222


def init_upsampler_params(in_channels: int = 128, mid_channels: int = 1024, num_blocks_per_stage: int = 4, seed: int = 0) -> Dict[str, Tensor]:
    """Seeded weights in the reference's layouts.  Convolutions U(+-1/sqrt(fan_in)) like upsampler.py:38-43; biases and
    the GroupNorm affine are given non-trivial values (the reference initialises them to 0 / 1) so they are exercised."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, Tensor] = {}

    def conv(name, o, i, k):  # k: kernel dims
        bound = 1.0 / math.sqrt(i * math.prod(k))
        out[name + ".weight"] = (torch.rand(o, *k, i, generator=g) * 2 - 1) * bound
        out[name + ".bias"] = (torch.rand(o, generator=g) * 2 - 1) * bound

    def norm(name, c):
        out[name + ".weight"] = 1 + 0.1 * torch.randn(c, generator=g)
        out[name + ".bias"] = 0.1 * torch.randn(c, generator=g)

    def block(name, c):
        conv(name + ".conv1", c, c, (3, 3, 3))
        norm(name + ".norm1", c)
        conv(name + ".conv2", c, c, (3, 3, 3))
        norm(name + ".norm2", c)

    conv("initial_conv", mid_channels, in_channels, (3, 3, 3))
    norm("initial_norm", mid_channels)
    for i in range(num_blocks_per_stage):
        block(f"res_blocks.{i}", mid_channels)
    conv("upsampler.conv", 4 * mid_channels, mid_channels, (3, 3))
    for i in range(num_blocks_per_stage):
        block(f"post_upsample_res_blocks.{i}", mid_channels)
    conv("final_conv", in_channels, mid_channels, (3, 3, 3))
    return out


def conv3d_cl(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """upsampler.py:51-72: x (N, D, H, W, C_in), w (C_out, 3, 3, 3, C_in), stride 1, padding 1 -> (N, D, H, W, C_out)."""
    y = F.conv3d(x.permute(0, 4, 1, 2, 3), w.permute(0, 4, 1, 2, 3), b, stride=1, padding=1)
    return y.permute(0, 2, 3, 4, 1).contiguous()


def conv2d_cl(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """nn.Conv2d(mid, 4 mid, 3, padding=1) (upsampler.py:149): x (N, H, W, C_in), w (C_out, 3, 3, C_in)."""
    y = F.conv2d(x.permute(0, 3, 1, 2), w.permute(0, 3, 1, 2), b, stride=1, padding=1)
    return y.permute(0, 2, 3, 1).contiguous()


def group_norm_cl(x: Tensor, w: Tensor, b: Tensor, num_groups: int = NUM_GROUPS, eps: float = GN_EPS) -> Tensor:
    """upsampler.py:85-114: fp32 statistics over (D*H*W, C/groups) per sample and group, population variance."""
    n, d, h, ww, c = x.shape
    t = x.float().reshape(n, d * h * ww, num_groups, c // num_groups)
    mean = t.mean(dim=(1, 3), keepdim=True)
    var = t.var(dim=(1, 3), keepdim=True, correction=0)
    t = ((t - mean) / torch.sqrt(var + eps)).reshape(n, d, h, ww, c)
    return (t * w.float() + b.float()).to(x.dtype)


def pixel_shuffle_cl(x: Tensor, r: int = 2) -> Tensor:
    """upsampler.py:124-139: (N, H, W, out_c*r*r) -> (N, H*r, W*r, out_c); channel index = (out_c, r_h, r_w) row-major."""
    n, h, w, c = x.shape
    out_c = c // (r * r)
    return x.reshape(n, h, w, out_c, r, r).permute(0, 1, 4, 2, 5, 3).reshape(n, h * r, w * r, out_c)


def res_block(p: Dict[str, Tensor], name: str, x: Tensor) -> Tensor:
    """upsampler.py:186-199: silu(norm2(conv2(silu(norm1(conv1 x)))) + x)."""
    y = conv3d_cl(x, p[name + ".conv1.weight"], p[name + ".conv1.bias"])
    y = F.silu(group_norm_cl(y, p[name + ".norm1.weight"], p[name + ".norm1.bias"]))
    y = conv3d_cl(y, p[name + ".conv2.weight"], p[name + ".conv2.bias"])
    y = group_norm_cl(y, p[name + ".norm2.weight"], p[name + ".norm2.bias"])
    return F.silu(y + x)


def latent_upsampler(p: Dict[str, Tensor], latent: Tensor) -> Tensor:
    """upsampler.py:232-294: (B, C, F, H, W) -> (B, C, F, 2H, 2W)."""
    n_blocks = 1 + max(int(k.split(".")[1]) for k in p if k.startswith("res_blocks."))
    x = latent.permute(0, 2, 3, 4, 1)
    x = conv3d_cl(x, p["initial_conv.weight"], p["initial_conv.bias"])
    x = F.silu(group_norm_cl(x, p["initial_norm.weight"], p["initial_norm.bias"]))
    for i in range(n_blocks):
        x = res_block(p, f"res_blocks.{i}", x)
    n, d, h, w, c = x.shape  # upsampler.py:156-174: frame by frame 2-D conv + pixel shuffle
    y = conv2d_cl(x.reshape(n * d, h, w, c), p["upsampler.conv.weight"], p["upsampler.conv.bias"])
    x = pixel_shuffle_cl(y).reshape(n, d, 2 * h, 2 * w, c)
    for i in range(n_blocks):
        x = res_block(p, f"post_upsample_res_blocks.{i}", x)
    x = conv3d_cl(x, p["final_conv.weight"], p["final_conv.bias"])
    return x.permute(0, 4, 1, 2, 3).contiguous()


def upsample_latents(latent: Tensor, p: Dict[str, Tensor], latent_mean: Tensor, latent_std: Tensor) -> Tensor:
    """upsampler.py:297-316: un-normalise with the VAE's per-channel statistics, upsample, re-normalise."""
    mean, std = latent_mean.reshape(1, -1, 1, 1, 1), latent_std.reshape(1, -1, 1, 1, 1)
    return (latent_upsampler(p, latent * std + mean) - mean) / std


def sanitize_upsampler_weights(raw: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """upsampler.py:346-365: upstream (PyTorch) conv layouts -> the channels-last ones above."""
    out = {}
    for k, v in raw.items():
        if "conv" in k and "weight" in k and v.dim() == 5:
            v = v.permute(0, 2, 3, 4, 1)
        if "conv" in k and "weight" in k and v.dim() == 4:
            v = v.permute(0, 2, 3, 1)
        out[k] = v
    return out
