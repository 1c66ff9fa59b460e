"""ORACLE — test infrastructure, not product code.  GROUNDWORK for SURVEY.md §8f row N4, second half: no product path
exists for it yet (round 1); this restatement and its golden are what that path will be held to.

CPU restatement (PyTorch fp32) of the reference's LTX-2 video VAE decoder forward, the step after the last denoise loop:
``mlx_video/models/ltx/video_vae/decoder.py`` — ``get_timestep_embedding`` :29-54, ``TimestepEmbedding`` /
``PixArtAlphaTimestepEmbedder`` :57-91, ``ResnetBlock3DSimple`` :94-180, ``ResBlockGroup`` :183-234,
``LTX2VideoDecoder.__call__`` :361-450 (non-tiled, non-chunked path); ``convolution.py`` — ``reflect_pad_2d`` :13-40,
``CausalConv3d.__call__`` :120-166; ``sampling.py`` — ``DepthToSpaceUpsample`` :106-197; ``ops.py`` — ``unpatchify`` :47-80.
Pinned by tests/golden/vae_decoder.npz, which oracle/make_golden_vae.py produced by running the reference's own
``LTX2VideoDecoder`` over the shim.  Tensors are channels-first (B, C, F, H, W) as in the reference; parameters use the
reference's names (``conv_in.conv.conv.weight`` (O, 3, 3, 3, I), ``up_blocks.{0,2,4,6}.res_blocks.{i}.conv{1,2}.conv.conv.*``,
``up_blocks.{..}.res_blocks.{i}.scale_shift_table`` (4, C), ``up_blocks.{..}.time_embedder.timestep_embedder.linear_{1,2}.*``,
``up_blocks.{1,3,5}.conv.conv.*``, ``conv_out.conv.conv.*``, ``last_time_embedder.*``, ``last_scale_shift_table`` (2, 128),
``latents_mean`` / ``latents_std`` (128,), ``timestep_scale_multiplier``).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
WIDTHS = (1024, 512, 256, 128)  # decoder.py:237-250: res-block groups up_blocks.{0,2,4,6}; upsamplers up_blocks.{1,3,5}
LAYERS_PER_BLOCK = 5
DECODE_NOISE_SCALE, DECODE_TIMESTEP = 0.025, 0.05  # decoder.py:267-268


def init_decoder_params(seed: int = 0, table_std: float = 0.05) -> Dict[str, Tensor]:
    """Seeded weights of the full-width decoder (545 M parameters).  Convolutions / linears U(+-1/sqrt(fan_in)); the
    scale-shift tables, which the reference initialises to zero, get small random values so the timestep path matters."""
    g = torch.Generator().manual_seed(seed)
    p: Dict[str, Tensor] = {}

    def conv(name, o, i):
        bound = 1.0 / math.sqrt(i * 27)
        p[name + ".weight"] = (torch.rand(o, 3, 3, 3, i, generator=g) * 2 - 1) * bound
        p[name + ".bias"] = (torch.rand(o, generator=g) * 2 - 1) * bound

    def linear(name, o, i):
        bound = 1.0 / math.sqrt(i)
        p[name + ".weight"] = (torch.rand(o, i, generator=g) * 2 - 1) * bound
        p[name + ".bias"] = (torch.rand(o, generator=g) * 2 - 1) * bound

    def embedder(name, dim):
        linear(name + ".timestep_embedder.linear_1", dim, 256)
        linear(name + ".timestep_embedder.linear_2", dim, dim)

    conv("conv_in.conv.conv", 1024, 128)
    for level, c in enumerate(WIDTHS):
        blk = f"up_blocks.{2 * level}"
        embedder(blk + ".time_embedder", 4 * c)
        for i in range(LAYERS_PER_BLOCK):
            conv(f"{blk}.res_blocks.{i}.conv1.conv.conv", c, c)
            conv(f"{blk}.res_blocks.{i}.conv2.conv.conv", c, c)
            p[f"{blk}.res_blocks.{i}.scale_shift_table"] = table_std * torch.randn(4, c, generator=g)
        if level < 3:
            conv(f"up_blocks.{2 * level + 1}.conv.conv", (c // 2) * 8, c)
    conv("conv_out.conv.conv", 48, 128)
    embedder("last_time_embedder", 256)
    p["last_scale_shift_table"] = table_std * torch.randn(2, 128, generator=g)
    p["latents_mean"] = 0.2 * torch.randn(128, generator=g)
    p["latents_std"] = 0.5 + torch.rand(128, generator=g)
    p["timestep_scale_multiplier"] = torch.tensor(1000.0)
    return p


def causal_conv3d(x: Tensor, w: Tensor, b: Tensor, causal: bool) -> Tensor:
    """convolution.py:120-166 (kernel 3, stride 1): temporal padding by frame replication — two copies of the first frame
    when causal, one of the first and one of the last otherwise — REFLECT padding of H and W, then a valid convolution.
    x (B, C, D, H, W); w in the reference layout (O, 3, 3, 3, I)."""
    if causal:
        x = torch.cat([x[:, :, :1].repeat(1, 1, 2, 1, 1), x], dim=2)
    else:
        x = torch.cat([x[:, :, :1], x, x[:, :, -1:]], dim=2)
    x = F.pad(x, (1, 1, 1, 1, 0, 0), mode="reflect")  # convolution.py:13-40: reflection excludes the border pixel
    return F.conv3d(x, w.permute(0, 4, 1, 2, 3), b)


def pixel_norm(x: Tensor, eps: float = 1e-8) -> Tensor:
    """decoder.py:136-138,357-359: x / sqrt(mean over channels of x^2 + eps)."""
    return x / torch.sqrt(torch.mean(x * x, dim=1, keepdim=True) + eps)


def timestep_embedding_256(t: Tensor) -> Tensor:
    """decoder.py:29-54 with embedding_dim 256, flip_sin_to_cos, downscale_freq_shift 0: [cos | sin](t * 10000^(-i/128))."""
    half = 128
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32) / half)
    arg = t.float()[:, None] * freqs[None, :]
    return torch.cat([torch.cos(arg), torch.sin(arg)], dim=-1)


def timestep_embedder(p: Dict[str, Tensor], name: str, t: Tensor) -> Tensor:
    """decoder.py:73-91: sinusoid -> linear_1 -> SiLU -> linear_2."""
    e = timestep_embedding_256(t)
    e = F.silu(F.linear(e, p[name + ".timestep_embedder.linear_1.weight"], p[name + ".timestep_embedder.linear_1.bias"]))
    return F.linear(e, p[name + ".timestep_embedder.linear_2.weight"], p[name + ".timestep_embedder.linear_2.bias"])


def resnet_block(p: Dict[str, Tensor], name: str, x: Tensor, causal: bool, ts_embed: Optional[Tensor]) -> Tensor:
    """decoder.py:140-180: pixel_norm -> (1 + scale1) x + shift1 -> SiLU -> conv1 -> pixel_norm -> (1 + scale2) x + shift2 ->
    SiLU -> conv2, plus the input.  Table rows: shift1, scale1, shift2, scale2."""
    B, C = x.shape[:2]
    h = pixel_norm(x)
    if ts_embed is not None:
        ada = p[name + ".scale_shift_table"][None, :, :, None, None, None] + ts_embed.reshape(B, 4, C, 1, 1, 1)
        shift1, scale1, shift2, scale2 = ada[:, 0], ada[:, 1], ada[:, 2], ada[:, 3]
        h = h * (1 + scale1) + shift1
    h = causal_conv3d(F.silu(h), p[name + ".conv1.conv.conv.weight"], p[name + ".conv1.conv.conv.bias"], causal)
    h = pixel_norm(h)
    if ts_embed is not None:
        h = h * (1 + scale2) + shift2
    h = causal_conv3d(F.silu(h), p[name + ".conv2.conv.conv.weight"], p[name + ".conv2.conv.conv.bias"], causal)
    return h + x


def depth_to_space(x: Tensor, st: int = 2, sh: int = 2, sw: int = 2) -> Tensor:
    """sampling.py:143-157: (B, C*st*sh*sw, D, H, W) -> (B, C, D*st, H*sh, W*sw), channel index = (c, st, sh, sw) row-major."""
    B, cp, D, H, W = x.shape
    c = cp // (st * sh * sw)
    return x.reshape(B, c, st, sh, sw, D, H, W).permute(0, 1, 5, 2, 6, 3, 7, 4).reshape(B, c, D * st, H * sh, W * sw)


def depth_to_space_upsample(p: Dict[str, Tensor], name: str, x: Tensor, causal: bool) -> Tensor:
    """sampling.py:159-197 (stride (2,2,2), residual, channel reduction 2): conv to 4x channels -> depth-to-space -> drop
    the first frame; residual = depth-to-space of the INPUT tiled (not element-repeated) 4x over channels, first frame
    dropped as well."""
    res = depth_to_space(x).repeat(1, 4, 1, 1, 1)[:, :, 1:]
    y = causal_conv3d(x, p[name + ".conv.conv.weight"], p[name + ".conv.conv.bias"], causal)
    return depth_to_space(y)[:, :, 1:] + res


def unpatchify(x: Tensor, patch: int = 4) -> Tensor:
    """ops.py:47-80 with patch_size_t = 1: channel index = (c, pt, pr (width), pq (height)) -> (B, C, F, H*patch, W*patch)."""
    B, cp, Fr, H, W = x.shape
    c = cp // (patch * patch)
    return x.reshape(B, c, 1, patch, patch, Fr, H, W).permute(0, 1, 5, 2, 6, 4, 7, 3).reshape(B, c, Fr, H * patch, W * patch)


def decode(p: Dict[str, Tensor], sample: Tensor, causal: bool = False, timestep: Optional[Tensor] = None,
           noise: Optional[Tensor] = None, noise_scale: float = DECODE_NOISE_SCALE) -> Tensor:
    """decoder.py:361-450: latents (B, 128, F, H, W) -> video (B, 3, 8(F-1)+1, 32H, 32W).  ``noise``: the N(0,1) draw the
    reference takes from mx.random.normal (pass it explicitly to reproduce a run; None = zeros)."""
    B = sample.shape[0]
    n = torch.zeros_like(sample) if noise is None else noise
    sample = n * noise_scale + (1.0 - noise_scale) * sample                       # decoder.py:380-382
    sample = sample * p["latents_std"].float().reshape(1, -1, 1, 1, 1) + p["latents_mean"].float().reshape(1, -1, 1, 1, 1)
    if timestep is None:
        timestep = torch.full((B,), DECODE_TIMESTEP)
    scaled = timestep * p["timestep_scale_multiplier"]
    x = causal_conv3d(sample, p["conv_in.conv.conv.weight"], p["conv_in.conv.conv.bias"], causal)
    for level in range(4):
        blk = f"up_blocks.{2 * level}"
        ts = timestep_embedder(p, blk + ".time_embedder", scaled.flatten())       # decoder.py:222-230
        for i in range(LAYERS_PER_BLOCK):
            x = resnet_block(p, f"{blk}.res_blocks.{i}", x, causal, ts)
        if level < 3:
            x = depth_to_space_upsample(p, f"up_blocks.{2 * level + 1}", x, causal)
    x = pixel_norm(x)
    e = timestep_embedder(p, "last_time_embedder", scaled.flatten()).reshape(B, 2, 128, 1, 1, 1)
    ada = p["last_scale_shift_table"][None, :, :, None, None, None] + e
    x = x * (1 + ada[:, 1]) + ada[:, 0]                                            # rows: shift, scale
    x = causal_conv3d(F.silu(x), p["conv_out.conv.conv.weight"], p["conv_out.conv.conv.bias"], causal)
    return unpatchify(x, 4)
