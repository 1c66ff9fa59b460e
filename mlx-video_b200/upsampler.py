"""Latent 2x spatial upsampler between the two stages of the LTX-2 pipelines (first half).

Mirror of the reference's ``mlx_video/models/ltx/upsampler.py``: ``LatentUpsampler(in_channels, mid_channels,
num_blocks_per_stage)`` :202-294 (initial Conv3d + GroupNorm + SiLU, ResBlock3D x n, frame-by-frame Conv2d + pixel shuffle,
ResBlock3D x n, final Conv3d), ``upsample_latents(latent, upsampler, latent_mean, latent_std)`` :297-316 and
``load_upsampler(weights_path)`` :319-373 — same names, argument meaning and parameter names
(``initial_conv.weight``, ``res_blocks.{i}.conv1.weight``, ``upsampler.conv.weight`` ...).

B200 layout: activations are channels-last fp32 (N, F, H, W, C) for the whole chain; every convolution is one tensor-core
GEMM (``ltxb_gemm_bf16``, fp32 accumulate and output, bias in the epilogue) over the rows ``ltxb_im2col_cl`` gathers in
bf16 — the reference's weight layout (C_out, kd, kh, kw, C_in) is the GEMM's W operand as stored, so weights are only
cast, never shuffled.  GroupNorm + affine (+ residual) + SiLU is one kernel pair (``ltxb_groupnorm_silu``); the
channels-first <-> channels-last moves carry the VAE un- / re-normalisation (``ltxb_latent_layout``).  No torch
arithmetic on the path; the module raises on CPU tensors like the rest of the package.
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, Iterator, Optional, Tuple, Union

import torch

from . import _lib, ops
from ._lib import LtxbError

Tensor = torch.Tensor
BF16, F32 = torch.bfloat16, torch.float32
NUM_GROUPS, GN_EPS = 32, 1e-5  # upsampler.py:77,182,184,222


class _Conv:
    """Conv3d (upsampler.py:6-72) / nn.Conv2d as a GEMM: weight bf16 [C_out, taps * C_in], bias f32 [C_out]."""

    def __init__(self, in_channels: int, out_channels: int, kernel: Tuple[int, int, int], device) -> None:
        self.in_channels, self.out_channels, self.kernel = in_channels, out_channels, kernel
        taps = kernel[0] * kernel[1] * kernel[2]
        self.weight = torch.zeros(out_channels, taps * in_channels, dtype=BF16, device=device)
        self.bias = torch.zeros(out_channels, dtype=F32, device=device)

    def __call__(self, x: Tensor, cols: Tensor) -> Tensor:
        """x f32 (N, D, H, W, C_in) -> f32 (N, D, H, W, C_out); ``cols`` is scratch for the gathered rows."""
        N, D, H, W, _ = x.shape
        M, K = N * D * H * W, self.weight.shape[1]
        a = cols[: M * K].view(M, K)
        ops.im2col_cl(x, a, *self.kernel)
        out = torch.empty(N, D, H, W, self.out_channels, dtype=F32, device=x.device)
        ops.gemm(a, self.weight, self.bias, out.view(M, self.out_channels), mode=_lib.EPI_BIAS_F32)
        return out


class _Norm:
    """GroupNorm3d(32, C) (upsampler.py:75-114)."""

    def __init__(self, channels: int, device) -> None:
        self.weight = torch.ones(channels, dtype=F32, device=device)
        self.bias = torch.zeros(channels, dtype=F32, device=device)

    def __call__(self, x: Tensor, resid: Optional[Tensor] = None, silu: bool = True) -> Tensor:
        return ops.groupnorm_silu(x, x, NUM_GROUPS, GN_EPS, self.weight, self.bias, resid=resid, silu=silu)  # in place


class ResBlock3D:
    """upsampler.py:177-199: silu(norm2(conv2(silu(norm1(conv1 x)))) + x)."""

    def __init__(self, channels: int, device) -> None:
        self.conv1, self.norm1 = _Conv(channels, channels, (3, 3, 3), device), _Norm(channels, device)
        self.conv2, self.norm2 = _Conv(channels, channels, (3, 3, 3), device), _Norm(channels, device)

    def __call__(self, x: Tensor, cols: Tensor) -> Tensor:
        y = self.norm1(self.conv1(x, cols))
        return self.norm2(self.conv2(y, cols), resid=x)  # the SiLU comes AFTER the residual add (upsampler.py:196-197)


class LatentUpsampler:
    """upsampler.py:202-294.  ``__call__(latent (B, C, F, H, W)) -> (B, C, F, 2H, 2W)``, f32 or bf16 in, same dtype out."""

    def __init__(self, in_channels: int = 128, mid_channels: int = 1024, num_blocks_per_stage: int = 4,
                 device: Union[str, torch.device, None] = None) -> None:
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        if device is None or torch.device(device).type != "cuda":
            raise LtxbError("LatentUpsampler needs a CUDA device (sm_100a); there is no CPU fallback on this path")
        if mid_channels % 64 or in_channels % 64 or mid_channels > 1024:  # GEMM K = taps * C in blocks of 64; GroupNorm smem
            raise ValueError(f"unsupported widths in={in_channels} mid={mid_channels} (multiples of 64, mid <= 1024)")
        self.device = dev = torch.device(device)
        self.in_channels, self.mid_channels = in_channels, mid_channels
        self.initial_conv = _Conv(in_channels, mid_channels, (3, 3, 3), dev)
        self.initial_norm = _Norm(mid_channels, dev)
        self.res_blocks: Dict[int, ResBlock3D] = {i: ResBlock3D(mid_channels, dev) for i in range(num_blocks_per_stage)}
        self.upsampler_conv = _Conv(mid_channels, 4 * mid_channels, (1, 3, 3), dev)  # SpatialRationalResampler.conv, per frame
        self.post_upsample_res_blocks: Dict[int, ResBlock3D] = {i: ResBlock3D(mid_channels, dev) for i in range(num_blocks_per_stage)}
        self.final_conv = _Conv(mid_channels, in_channels, (3, 3, 3), dev)
        self._cols: Optional[Tensor] = None

    # ------------------------------------------------------------------ parameters (reference names, reference layouts)
    def _convs(self) -> Iterator[Tuple[str, _Conv]]:
        yield "initial_conv", self.initial_conv
        for tag, blocks in (("res_blocks", self.res_blocks), ("post_upsample_res_blocks", self.post_upsample_res_blocks)):
            for i, b in blocks.items():
                yield f"{tag}.{i}.conv1", b.conv1
                yield f"{tag}.{i}.conv2", b.conv2
        yield "upsampler.conv", self.upsampler_conv
        yield "final_conv", self.final_conv

    def _norms(self) -> Iterator[Tuple[str, _Norm]]:
        yield "initial_norm", self.initial_norm
        for tag, blocks in (("res_blocks", self.res_blocks), ("post_upsample_res_blocks", self.post_upsample_res_blocks)):
            for i, b in blocks.items():
                yield f"{tag}.{i}.norm1", b.norm1
                yield f"{tag}.{i}.norm2", b.norm2

    def parameter_names(self):
        return [f"{n}.{s}" for n, _ in list(self._convs()) + list(self._norms()) for s in ("weight", "bias")]

    def load_weights(self, weights, strict: bool = True) -> None:
        """Weights in the reference's (MLX) layouts: conv3d (C_out, 3, 3, 3, C_in), conv2d (C_out, 3, 3, C_in) — flattened
        they ARE the GEMM operand.  Unknown names are ignored unless ``strict`` (the reference loads with strict=False and
        upstream files carry ``upsampler.blur_down.kernel``, upsampler.py:358-368)."""
        weights = dict(weights)
        names = set(self.parameter_names())
        missing = sorted(names - set(weights))
        extra = sorted(k for k in weights if k not in names)
        if strict and (missing or extra):
            raise ValueError(f"upsampler weights: missing {missing[:6]} unexpected {extra[:6]}")
        for name, conv in self._convs():
            w = weights.get(name + ".weight")
            if w is not None:
                if w.shape[0] != conv.out_channels or w.numel() != conv.weight.numel() or w.shape[-1] != conv.in_channels:
                    raise ValueError(f"shape mismatch for {name}.weight: {tuple(w.shape)}")
                conv.weight.copy_(w.reshape(conv.out_channels, -1).to(device=self.device, dtype=BF16))
            if weights.get(name + ".bias") is not None:
                conv.bias.copy_(weights[name + ".bias"].to(device=self.device, dtype=F32))
        for name, norm in self._norms():
            for s in ("weight", "bias"):
                if weights.get(f"{name}.{s}") is not None:
                    getattr(norm, s).copy_(weights[f"{name}.{s}"].to(device=self.device, dtype=F32))

    # ------------------------------------------------------------------ forward
    def _scratch(self, elems: int) -> Tensor:
        if self._cols is None or self._cols.numel() < elems:
            self._cols = torch.empty(elems, dtype=BF16, device=self.device)
        return self._cols

    def _forward(self, latent: Tensor, mean: Optional[Tensor], std: Optional[Tensor]) -> Tensor:
        if not latent.is_cuda:
            raise LtxbError("LatentUpsampler takes CUDA tensors; there is no CPU fallback on this path")
        if latent.dim() != 5 or latent.shape[1] != self.in_channels:
            raise ValueError(f"latent of shape {tuple(latent.shape)}: expected (B, {self.in_channels}, F, H, W)")
        B, C, F_, H, W = latent.shape
        out_dtype = latent.dtype
        x_cf = latent.to(F32).contiguous()  # plumbing: storage cast of a (B, 128, F, H, W) latent
        S = F_ * H * W
        cols = self._scratch(B * F_ * 4 * H * W * 27 * self.mid_channels)  # the widest operand: a post-upsample 3x3x3 conv
        x = torch.empty(B, F_, H, W, C, dtype=F32, device=self.device)
        ops.latent_layout(x_cf.view(B, C, S), x.view(B, S, C), std, mean, True)  # -> channels last, latent * std + mean
        x = self.initial_norm(self.initial_conv(x, cols))
        for i in sorted(self.res_blocks):
            x = self.res_blocks[i](x, cols)
        # SpatialRationalResampler (upsampler.py:156-174): 3x3 conv per frame -> pixel shuffle
        y = self.upsampler_conv(x.view(B * F_, 1, H, W, self.mid_channels), cols)
        x = torch.empty(B, F_, 2 * H, 2 * W, self.mid_channels, dtype=F32, device=self.device)
        ops.pixel_shuffle2(y.view(B * F_, H, W, 4 * self.mid_channels), x.view(B * F_, 2 * H, 2 * W, self.mid_channels))
        for i in sorted(self.post_upsample_res_blocks):
            x = self.post_upsample_res_blocks[i](x, cols)
        x = self.final_conv(x, cols)
        out = torch.empty(B, C, F_, 2 * H, 2 * W, dtype=F32, device=self.device)
        ops.latent_layout(x.view(B, 4 * S, C), out.view(B, C, 4 * S), std, mean, False)  # -> channels first, (x - mean) / std
        return out if out_dtype == F32 else out.to(out_dtype)

    def __call__(self, latent: Tensor, debug: bool = False) -> Tensor:
        return self._forward(latent, None, None)


def upsample_latents(latent: Tensor, upsampler: LatentUpsampler, latent_mean: Tensor, latent_std: Tensor,
                     debug: bool = False) -> Tensor:
    """upsampler.py:297-316: latent * std + mean -> upsampler -> (x - mean) / std, the (un)normalisation folded into the
    two layout kernels."""
    dev = upsampler.device
    mean = latent_mean.reshape(-1).to(device=dev, dtype=F32).contiguous()
    std = latent_std.reshape(-1).to(device=dev, dtype=F32).contiguous()
    if mean.numel() != upsampler.in_channels or std.numel() != upsampler.in_channels:
        raise ValueError("latent_mean / latent_std must have one value per latent channel")
    return upsampler._forward(latent, mean, std)


def load_upsampler(weights_path: Union[str, Path], device=None) -> LatentUpsampler:
    """upsampler.py:319-373: upstream safetensors (PyTorch conv layouts (O, I, D, H, W) / (O, I, H, W)) -> model; the width
    is read off ``res_blocks.0.conv1.weight``, four blocks per stage."""
    from safetensors import safe_open

    raw: Dict[str, Tensor] = {}
    with safe_open(str(weights_path), framework="pt", device="cpu") as f:
        for k in f.keys():
            raw[k] = f.get_tensor(k)
    sample = raw.get("res_blocks.0.conv1.weight")
    mid = int(sample.shape[0]) if sample is not None else 1024
    model = LatentUpsampler(in_channels=128, mid_channels=mid, num_blocks_per_stage=4, device=device)
    sanitized = {}
    for k, v in raw.items():
        if "conv" in k and "weight" in k and v.dim() == 5:
            v = v.permute(0, 2, 3, 4, 1)  # -> (O, D, H, W, I): layout plumbing at load time
        elif "conv" in k and "weight" in k and v.dim() == 4:
            v = v.permute(0, 2, 3, 1)
        sanitized[k] = v.contiguous()
    model.load_weights(sanitized, strict=False)
    return model


__all__ = ["LatentUpsampler", "ResBlock3D", "upsample_latents", "load_upsampler"]
