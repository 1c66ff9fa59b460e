"""LTX-2 video VAE decoder on sm_100a kernels — SURVEY.md section 8f row N4, second half (the step after the last denoise
loop).

Host-side mirror of the reference's ``mlx_video/models/ltx/video_vae/decoder.py`` (``LTX2VideoDecoder`` :237-450,
``decode_tiled`` :452-531) and ``tiling.py`` (``TilingConfig`` :97-229, ``decode_with_tiling`` :279-520): same class and
function names, argument names and meaning, parameter names (``conv_in.conv.conv.weight`` (O, 3, 3, 3, I),
``up_blocks.{0,2,4,6}.res_blocks.{i}.conv{1,2}.conv.conv.*``, ``...scale_shift_table`` (4, C),
``up_blocks.{..}.time_embedder.timestep_embedder.linear_{1,2}.*``, ``up_blocks.{1,3,5}.conv.conv.*``, ``conv_out.conv.conv.*``,
``last_time_embedder.*``, ``last_scale_shift_table``, ``latents_mean`` / ``latents_std``, ``timestep_scale_multiplier``).

Data layout: activations stay channels-last fp32 [B, F, H, W, C]; every CausalConv3d is ONE tcgen05 GEMM
(``ltxb_gemm_bf16``, fp32 out, bias — and the ResNet skip — in the epilogue) over bf16 rows gathered by
``ltxb_vae_gather_rows``, which applies replicate-in-time / reflect-in-space padding and the pixel-norm + AdaLN + SiLU chain
in front of the convolution while it gathers.  The reference's weight layout (O, 3, 3, 3, I) IS the GEMM's W [N, K]
operand: weights are cast to bf16, never shuffled.  Rows are produced in chunks so the gathered operand stays bounded.
There is no CPU path and no torch arithmetic on the path (torch supplies memory and the stream).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib, ops
from ._lib import LtxbError

Tensor = torch.Tensor
BF16, F32 = torch.bfloat16, torch.float32
WIDTHS = (1024, 512, 256, 128)  # decoder.py:296-322: res-block groups up_blocks.{0,2,4,6}; upsamplers up_blocks.{1,3,5}


# ------------------------------------------------------------------------------------------------ tiling.py
def compute_trapezoidal_mask_1d(length: int, ramp_left: int, ramp_right: int, left_starts_from_0: bool = False) -> np.ndarray:
    """tiling.py:17-60 — ones with a linear fade-in over ``ramp_left`` samples (from 0 when ``left_starts_from_0``, else from
    the first non-zero step) and a linear fade-out over ``ramp_right`` samples that stops short of 0.  fp32, like the
    reference's list arithmetic rounded into an mx.array."""
    if length <= 0:
        raise ValueError("Mask length must be positive.")
    ramp_left, ramp_right = max(0, min(ramp_left, length)), max(0, min(ramp_right, length))
    mask = np.ones(length, dtype=np.float64)
    if ramp_left > 0:
        first = 0 if left_starts_from_0 else 1
        steps = ramp_left + (0 if left_starts_from_0 else 1)  # = interval_length - 1
        mask[:ramp_left] *= (np.arange(ramp_left) + first) / steps
    if ramp_right > 0:
        mask[length - ramp_right:] *= (ramp_right - np.arange(ramp_right)) / (ramp_right + 1)
    return np.clip(mask, 0.0, 1.0).astype(np.float32)


@dataclass(frozen=True)
class SpatialTilingConfig:
    """tiling.py:63-80"""

    tile_size_in_pixels: int
    tile_overlap_in_pixels: int = 0

    def __post_init__(self) -> None:
        s, o = self.tile_size_in_pixels, self.tile_overlap_in_pixels
        if s < 64:
            raise ValueError(f"tile_size_in_pixels must be at least 64, got {s}")
        if s % 32 != 0:
            raise ValueError(f"tile_size_in_pixels must be divisible by 32, got {s}")
        if o % 32 != 0:
            raise ValueError(f"tile_overlap_in_pixels must be divisible by 32, got {o}")
        if o >= s:
            raise ValueError(f"Overlap must be less than tile size, got {o} and {s}")


@dataclass(frozen=True)
class TemporalTilingConfig:
    """tiling.py:83-100"""

    tile_size_in_frames: int
    tile_overlap_in_frames: int = 0

    def __post_init__(self) -> None:
        s, o = self.tile_size_in_frames, self.tile_overlap_in_frames
        if s < 16:
            raise ValueError(f"tile_size_in_frames must be at least 16, got {s}")
        if s % 8 != 0:
            raise ValueError(f"tile_size_in_frames must be divisible by 8, got {s}")
        if o % 8 != 0:
            raise ValueError(f"tile_overlap_in_frames must be divisible by 8, got {o}")
        if o >= s:
            raise ValueError(f"Overlap must be less than tile size, got {o} and {s}")


@dataclass(frozen=True)
class TilingConfig:
    """tiling.py:103-229 — spatial / temporal tile sizes with overlap; the named presets of the reference."""

    spatial_config: Optional[SpatialTilingConfig] = None
    temporal_config: Optional[TemporalTilingConfig] = None

    @classmethod
    def default(cls) -> "TilingConfig":
        return cls(SpatialTilingConfig(512, 64), TemporalTilingConfig(64, 24))

    @classmethod
    def spatial_only(cls, tile_size: int = 512, overlap: int = 64) -> "TilingConfig":
        return cls(SpatialTilingConfig(tile_size, overlap), None)

    @classmethod
    def temporal_only(cls, tile_size: int = 64, overlap: int = 24) -> "TilingConfig":
        return cls(None, TemporalTilingConfig(tile_size, overlap))

    @classmethod
    def aggressive(cls) -> "TilingConfig":
        return cls(SpatialTilingConfig(256, 64), TemporalTilingConfig(32, 8))

    @classmethod
    def conservative(cls) -> "TilingConfig":
        return cls(SpatialTilingConfig(768, 64), TemporalTilingConfig(96, 24))

    @classmethod
    def auto(cls, height: int, width: int, num_frames: int, spatial_threshold: int = 512,
             temporal_threshold: int = 65) -> Optional["TilingConfig"]:
        """tiling.py:152-229: no tiling below the thresholds; aggressive above 2 GB of fp32 output (or > 768x1024 and > 100
        frames); otherwise tile sizes by resolution / frame count."""
        spatial, temporal = max(height, width) > spatial_threshold, num_frames > temporal_threshold
        if not spatial and not temporal:
            return None
        if (3 * num_frames * height * width * 4) / 1024 ** 3 > 2.0 or (height * width > 768 * 1024 and num_frames > 100):
            return cls.aggressive()
        s_cfg = t_cfg = None
        if spatial:
            s_cfg = SpatialTilingConfig(512 if 768 < max(height, width) <= 1024 else 384, 64)
        if temporal:
            size, overlap = (32, 8) if num_frames > 200 else ((48, 16) if num_frames > 100 else (64, 24))
            t_cfg = TemporalTilingConfig(size, overlap)
        return cls(s_cfg, t_cfg)


@dataclass
class DimensionIntervals:
    """tiling.py:232-238"""

    starts: List[int]
    ends: List[int]
    left_ramps: List[int]
    right_ramps: List[int]


def split_in_spatial(size: int, overlap: int, dimension_size: int) -> DimensionIntervals:
    """tiling.py:241-253: tiles of ``size`` every ``size - overlap``, the last one cut at the border; ramps = the overlaps."""
    if dimension_size <= size:
        return DimensionIntervals([0], [dimension_size], [0], [0])
    step = size - overlap
    n = (dimension_size + size - 2 * overlap - 1) // step
    starts = [i * step for i in range(n)]
    ends = [s + size for s in starts[:-1]] + [dimension_size]
    return DimensionIntervals(starts, ends, [0] + [overlap] * (n - 1), [overlap] * (n - 1) + [0])


def split_in_temporal(size: int, overlap: int, dimension_size: int) -> DimensionIntervals:
    """tiling.py:256-273: the spatial split with every tile but the first starting one latent frame earlier (the causal
    first frame of a tile decodes to ONE output frame) and a left ramp one longer."""
    iv = split_in_spatial(size, overlap, dimension_size)
    if len(iv.starts) == 1:
        return iv
    return DimensionIntervals([iv.starts[0]] + [s - 1 for s in iv.starts[1:]], iv.ends,
                              [iv.left_ramps[0]] + [r + 1 for r in iv.left_ramps[1:]], iv.right_ramps)


def map_temporal_slice(begin: int, end: int, left_ramp: int, right_ramp: int, scale: int) -> Tuple[slice, np.ndarray]:
    """tiling.py:276-285: latent frames [begin, end) -> output frames [begin*scale, 1 + (end-1)*scale) and their mask."""
    start, stop = begin * scale, 1 + (end - 1) * scale
    left = 1 + (left_ramp - 1) * scale if left_ramp > 0 else 0
    return slice(start, stop), compute_trapezoidal_mask_1d(stop - start, left, right_ramp * scale, True)


def map_spatial_slice(begin: int, end: int, left_ramp: int, right_ramp: int, scale: int) -> Tuple[slice, np.ndarray]:
    """tiling.py:288-296"""
    start, stop = begin * scale, end * scale
    return slice(start, stop), compute_trapezoidal_mask_1d(stop - start, left_ramp * scale, right_ramp * scale, False)


def decode_with_tiling(decoder_fn: Callable, latents: Tensor, tiling_config: TilingConfig, spatial_scale: int = 32,
                       temporal_scale: int = 8, causal: bool = False, timestep: Optional[Tensor] = None,
                       chunked_conv: bool = False, on_frames_ready: Optional[Callable[[Tensor, int], None]] = None) -> Tensor:
    """tiling.py:299-520: decode overlapping (time, height, width) tiles of ``latents`` (B, C, F, H, W) and blend them with
    separable trapezoid masks — accumulate ``tile * mask`` and ``mask`` (``ltxb_vae_blend_tile``), divide at the end
    (``ltxb_vae_blend_normalize``).  ``on_frames_ready(frames, start)`` is called with the frames no later tile can touch
    after each temporal tile, like the reference's streaming hook."""
    if not latents.is_cuda:
        raise LtxbError("decode_with_tiling needs CUDA latents; there is no CPU fallback on this path")
    b, _, f_lat, h_lat, w_lat = latents.shape
    out_f, out_h, out_w = 1 + (f_lat - 1) * temporal_scale, h_lat * spatial_scale, w_lat * spatial_scale
    s_cfg, t_cfg = tiling_config.spatial_config, tiling_config.temporal_config
    s_tile, s_ov = ((s_cfg.tile_size_in_pixels // spatial_scale, s_cfg.tile_overlap_in_pixels // spatial_scale)
                    if s_cfg is not None else (max(h_lat, w_lat), 0))
    t_tile, t_ov = ((t_cfg.tile_size_in_frames // temporal_scale, t_cfg.tile_overlap_in_frames // temporal_scale)
                    if t_cfg is not None else (f_lat, 0))
    t_iv = split_in_temporal(t_tile, t_ov, f_lat)
    h_iv, w_iv = split_in_spatial(s_tile, s_ov, h_lat), split_in_spatial(s_tile, s_ov, w_lat)
    dev = latents.device
    output = torch.zeros(b, 3, out_f, out_h, out_w, dtype=F32, device=dev)
    weights = torch.zeros(b, 1, out_f, out_h, out_w, dtype=F32, device=dev)
    emitted = 0

    def finalized(lo: int, hi: int) -> Tensor:
        part, wpart = output[:, :, lo:hi].contiguous(), weights[:, :, lo:hi].contiguous()
        return ops.vae_blend_normalize(part, wpart).to(latents.dtype)

    for ti in range(len(t_iv.starts)):
        t_sl, t_mask = map_temporal_slice(t_iv.starts[ti], t_iv.ends[ti], t_iv.left_ramps[ti], t_iv.right_ramps[ti], temporal_scale)
        for hi in range(len(h_iv.starts)):
            h_sl, h_mask = map_spatial_slice(h_iv.starts[hi], h_iv.ends[hi], h_iv.left_ramps[hi], h_iv.right_ramps[hi], spatial_scale)
            for wi in range(len(w_iv.starts)):
                w_sl, w_mask = map_spatial_slice(w_iv.starts[wi], w_iv.ends[wi], w_iv.left_ramps[wi], w_iv.right_ramps[wi], spatial_scale)
                tile_lat = latents[:, :, t_iv.starts[ti]:t_iv.ends[ti], h_iv.starts[hi]:h_iv.ends[hi], w_iv.starts[wi]:w_iv.ends[wi]].contiguous()
                tile = decoder_fn(tile_lat, causal=causal, timestep=timestep, debug=False, chunked_conv=chunked_conv)
                tile = tile.to(F32).contiguous()
                at = min(tile.shape[2], t_sl.stop - t_sl.start)
                ah = min(tile.shape[3], h_sl.stop - h_sl.start)
                aw = min(tile.shape[4], w_sl.stop - w_sl.start)
                masks = [torch.from_numpy(np.ascontiguousarray(m)).to(dev) for m in (t_mask, h_mask, w_mask)]
                ops.vae_blend_tile(tile, at, ah, aw, masks[0], masks[1], masks[2], output, weights, t_sl.start, h_sl.start, w_sl.start)
        if on_frames_ready is not None and len(t_iv.starts) > 1 and ti < len(t_iv.starts) - 1:
            nxt = t_iv.starts[ti + 1]
            nxt_out = 0 if nxt == 0 else 1 + (nxt - 1) * temporal_scale  # first output frame the next tile contributes to
            if nxt_out > emitted:
                on_frames_ready(finalized(emitted, nxt_out), emitted)
                emitted = nxt_out
    ops.vae_blend_normalize(output, weights)
    if on_frames_ready is not None and emitted < out_f:
        on_frames_ready(output[:, :, emitted:].to(latents.dtype), emitted)
    return output.to(latents.dtype)


# ------------------------------------------------------------------------------------------------ decoder.py
class _Conv:
    """CausalConv3d parameters (convolution.py:43-118): weight bf16 [O, 27*I] (the reference layout (O, 3, 3, 3, I) flattened —
    the GEMM's W operand as stored), bias f32 [O]."""

    def __init__(self, c_out: int, c_in: int, device) -> None:
        self.c_in, self.c_out = c_in, c_out
        self.weight = torch.zeros(c_out, 27 * c_in, dtype=BF16, device=device)
        self.bias = torch.zeros(c_out, dtype=F32, device=device)


class _Linear:
    def __init__(self, n_out: int, n_in: int, device) -> None:
        self.weight = torch.zeros(n_out, n_in, dtype=BF16, device=device)
        self.bias = torch.zeros(n_out, dtype=F32, device=device)


class LTX2VideoDecoder:
    """decoder.py:237-531.  ``__call__(sample (B, 128, F, H, W), causal=False, timestep=None)`` -> video
    (B, 3, 8(F-1)+1, 32H, 32W) in ``sample``'s dtype.  ``noise``: the N(0,1) draw the reference takes from
    ``mx.random.normal`` (decoder.py:380-382) — pass it to reproduce a run; None draws it on the device."""

    # rows of the materialised convolution operand per gather / GEMM chunk (bounds the bf16 operand: 128 Ki x 27 C x 2 B)
    max_rows_per_chunk = 1 << 17

    def __init__(self, in_channels: int = 128, out_channels: int = 3, patch_size: int = 4, num_layers_per_block: int = 5,
                 spatial_padding_mode: str = "reflect", timestep_conditioning: bool = True, device=None) -> None:
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        if device is None or torch.device(device).type != "cuda":
            raise LtxbError("LTX2VideoDecoder needs a CUDA device (sm_100a); there is no CPU fallback on this path")
        if str(getattr(spatial_padding_mode, "value", spatial_padding_mode)).lower() != "reflect":
            raise LtxbError("only the reflect spatial padding the LTX-2 decoder is built with has a kernel")
        if patch_size != 4 or out_channels != 3:
            raise LtxbError("the un-patchify kernel is built for patch_size 4, 3 output channels (the LTX-2 decoder)")
        self.device = torch.device(device)
        self.patch_size, self.in_channels, self.timestep_conditioning = patch_size, in_channels, timestep_conditioning
        self.num_layers_per_block = num_layers_per_block
        self.decode_noise_scale, self.decode_timestep = 0.025, 0.05  # decoder.py:267-268
        dev = self.device
        self.latents_mean = torch.zeros(in_channels, dtype=F32, device=dev)
        self.latents_std = torch.ones(in_channels, dtype=F32, device=dev)
        self.timestep_scale_multiplier = torch.tensor(1000.0)
        self._p: Dict[str, Tensor] = {}
        self._convs: Dict[str, _Conv] = {}
        self._linears: Dict[str, _Linear] = {}

        def conv(name, o, i):
            c = self._convs[name] = _Conv(o, i, dev)
            self._p[name + ".weight"], self._p[name + ".bias"] = c.weight, c.bias

        def embedder(name, dim):
            for lin, (o, i) in (("linear_1", (dim, 256)), ("linear_2", (dim, dim))):
                m = self._linears[f"{name}.timestep_embedder.{lin}"] = _Linear(o, i, dev)
                self._p[f"{name}.timestep_embedder.{lin}.weight"], self._p[f"{name}.timestep_embedder.{lin}.bias"] = m.weight, m.bias

        conv("conv_in.conv.conv", WIDTHS[0], in_channels)
        for level, c in enumerate(WIDTHS):
            blk = f"up_blocks.{2 * level}"
            if timestep_conditioning:
                embedder(blk + ".time_embedder", 4 * c)
            for i in range(num_layers_per_block):
                conv(f"{blk}.res_blocks.{i}.conv1.conv.conv", c, c)
                conv(f"{blk}.res_blocks.{i}.conv2.conv.conv", c, c)
                if timestep_conditioning:
                    self._p[f"{blk}.res_blocks.{i}.scale_shift_table"] = torch.zeros(4, c, dtype=F32, device=dev)
            if level < 3:
                conv(f"up_blocks.{2 * level + 1}.conv.conv", (c // 2) * 8, c)
        conv("conv_out.conv.conv", out_channels * patch_size * patch_size, WIDTHS[-1])
        if timestep_conditioning:
            embedder("last_time_embedder", 2 * WIDTHS[-1])
            self._p["last_scale_shift_table"] = torch.zeros(2, WIDTHS[-1], dtype=F32, device=dev)
        self._ws: Dict[tuple, Tensor] = {}

    # ------------------------------------------------------------------ parameters
    def parameters(self) -> Dict[str, Tensor]:
        d = dict(self._p)
        d["latents_mean"], d["latents_std"], d["timestep_scale_multiplier"] = self.latents_mean, self.latents_std, self.timestep_scale_multiplier
        return d

    def load_weights(self, weights: Dict[str, Tensor], strict: bool = True) -> None:
        """Reference parameter names; convolution weights in the reference layout (O, 3, 3, 3, I) — or PyTorch's
        (O, I, 3, 3, 3), which load_vae_decoder transposes (decoder.py:575-578) — are flattened to the GEMM operand."""
        params = self.parameters()
        missing = [k for k in params if k not in weights]
        if strict and missing:
            raise ValueError(f"Missing {len(missing)} parameters: {missing[:8]}{'...' if len(missing) > 8 else ''}")
        for name, dst in params.items():
            src = weights.get(name)
            if src is None:
                continue
            if name == "timestep_scale_multiplier":
                self.timestep_scale_multiplier = torch.as_tensor(src, dtype=F32).reshape(())
                continue
            if src.dim() == 5:
                conv = self._convs[name[: -len(".weight")]]
                if tuple(src.shape) == (conv.c_out, conv.c_in, 3, 3, 3) and conv.c_in != 3:
                    src = src.permute(0, 2, 3, 4, 1)
                if tuple(src.shape) != (conv.c_out, 3, 3, 3, conv.c_in):
                    raise ValueError(f"shape mismatch for {name}: checkpoint {tuple(src.shape)}")
                src = src.reshape(conv.c_out, 27 * conv.c_in)
            if tuple(src.shape) != tuple(dst.shape):
                raise ValueError(f"shape mismatch for {name}: checkpoint {tuple(src.shape)} vs model {tuple(dst.shape)}")
            dst.copy_(src.to(device=dst.device, dtype=dst.dtype))  # plumbing: H2D copy + storage cast

    # ------------------------------------------------------------------ pieces
    def _buf(self, tag: str, shape, dtype) -> Tensor:
        key = (tag, tuple(int(s) for s in shape), dtype)
        t = self._ws.get(key)
        if t is None:
            t = self._ws[key] = torch.empty(key[1], dtype=dtype, device=self.device)
        return t

    def _conv(self, name: str, x: Tensor, causal: bool, out: Tensor, pre: Optional[dict] = None, resid: Optional[Tensor] = None) -> Tensor:
        """CausalConv3d of x f32 (B, D, H, W, C) into out f32 (B, D, H, W, O) [+ resid]; ``pre`` = the fused pre-op
        (pixel norm [+ AdaLN] + SiLU) arguments of ltxb_vae_gather_rows."""
        conv = self._convs[name]
        B, D, H, W, C = x.shape
        M = B * D * H * W
        chunk = min(M, self.max_rows_per_chunk)
        rows_buf = self._buf("rows", (chunk, 27 * C), BF16)
        out2 = out.view(M, conv.c_out)
        res2 = None if resid is None else resid.view(M, conv.c_out)
        for m0 in range(0, M, chunk):
            rows = min(chunk, M - m0)
            ops.vae_gather_rows(x, rows_buf, causal, m0, rows, pre_op=pre is not None, **(pre or {}))
            if res2 is None:
                ops.gemm(rows_buf[:rows], conv.weight, conv.bias, out2[m0:m0 + rows], _lib.EPI_BIAS_F32)
            else:
                ops.gemm(rows_buf[:rows], conv.weight, conv.bias, out2[m0:m0 + rows], _lib.EPI_RESID_GATE_F32, resid=res2[m0:m0 + rows])
        return out

    def _timestep_embed(self, name: str, scaled: Tensor) -> Tensor:
        """decoder.py:57-91: sinusoid(256) -> linear_1 -> SiLU -> linear_2, f32 [B, dim]."""
        B = scaled.numel()
        feat = torch.empty(B, 256, dtype=BF16, device=self.device)
        ops.timestep_embed(scaled.reshape(-1), 1.0, 256, feat)
        l1, l2 = self._linears[name + ".timestep_embedder.linear_1"], self._linears[name + ".timestep_embedder.linear_2"]
        h = torch.empty(B, l1.weight.shape[0], dtype=BF16, device=self.device)
        ops.gemm(feat, l1.weight, l1.bias, h, _lib.EPI_SILU_BF16)
        e = torch.empty(B, l2.weight.shape[0], dtype=F32, device=self.device)
        ops.gemm(h, l2.weight, l2.bias, e, _lib.EPI_BIAS_F32)
        return e

    # ------------------------------------------------------------------ forward (decoder.py:361-450)
    def __call__(self, sample: Tensor, causal: bool = False, timestep: Optional[Tensor] = None, debug: bool = False,
                 chunked_conv: bool = False, noise: Optional[Tensor] = None) -> Tensor:
        if not sample.is_cuda:
            raise LtxbError("LTX2VideoDecoder inputs must be CUDA tensors; there is no CPU fallback on this path")
        if sample.dim() != 5 or sample.shape[1] != self.in_channels:
            raise ValueError(f"latents must be (B, {self.in_channels}, F, H, W), got {tuple(sample.shape)}")
        B, C, F_, H, W = sample.shape
        dev = self.device
        x_in = sample.to(F32).contiguous()
        ns = float(self.decode_noise_scale) if self.timestep_conditioning else 0.0
        if ns != 0.0 and noise is None:
            noise = torch.randn(x_in.shape, dtype=F32, device=dev)  # the reference's mx.random.normal draw (decoder.py:381)
        x = torch.empty(B, F_, H, W, C, dtype=F32, device=dev)
        ops.vae_prepare_latent(x_in, None if ns == 0.0 else noise.to(F32).contiguous(), ns, self.latents_std, self.latents_mean, x)
        scaled = None
        if self.timestep_conditioning:
            if timestep is None:
                ts = torch.full((B,), float(self.decode_timestep), dtype=F32, device=dev)
            else:
                ts = timestep.to(device=dev, dtype=F32).reshape(-1)
            scaled = torch.empty_like(ts)
            scaled.copy_(ts)
            # decoder.py:396: timestep * timestep_scale_multiplier, folded into the sinusoid's scale argument below
        mult = float(self.timestep_scale_multiplier)
        y = torch.empty(B, F_, H, W, WIDTHS[0], dtype=F32, device=dev)
        self._conv("conv_in.conv.conv", x, causal, y)
        x = y
        for level, c in enumerate(WIDTHS):
            blk = f"up_blocks.{2 * level}"
            emb = None
            if self.timestep_conditioning:
                emb = self._timestep_embed_scaled(blk + ".time_embedder", scaled, mult)  # f32 [B, 4c]: shift1 | scale1 | shift2 | scale2
            h = torch.empty_like(x)
            for i in range(self.num_layers_per_block):
                name = f"{blk}.res_blocks.{i}"
                pre1, pre2 = {}, {}
                if emb is not None:
                    tab = self._p[name + ".scale_shift_table"]
                    pre1 = dict(table_shift=tab[0], table_scale=tab[1], emb_shift=emb[:, 0:c], emb_scale=emb[:, c:2 * c], emb_ld=4 * c)
                    pre2 = dict(table_shift=tab[2], table_scale=tab[3], emb_shift=emb[:, 2 * c:3 * c], emb_scale=emb[:, 3 * c:4 * c], emb_ld=4 * c)
                self._conv(name + ".conv1.conv.conv", x, causal, h, pre=pre1)
                self._conv(name + ".conv2.conv.conv", h, causal, x, pre=pre2, resid=x)  # x += conv2(...) in the GEMM epilogue
            if level < 3:
                Bx, D, Hh, Ww, _ = x.shape
                yc = torch.empty(Bx, D, Hh, Ww, 4 * c, dtype=F32, device=dev)
                self._conv(f"up_blocks.{2 * level + 1}.conv.conv", x, causal, yc)
                up = torch.empty(Bx, 2 * D - 1, 2 * Hh, 2 * Ww, c // 2, dtype=F32, device=dev)
                ops.vae_depth_to_space(yc, x, up)
                x = up
                del yc, h
        pre = {}
        if self.timestep_conditioning:
            e = self._timestep_embed_scaled("last_time_embedder", scaled, mult)  # [B, 2*128]: shift | scale
            tab = self._p["last_scale_shift_table"]
            pre = dict(table_shift=tab[0], table_scale=tab[1], emb_shift=e[:, 0:128], emb_scale=e[:, 128:256], emb_ld=256)
        Bx, D, Hh, Ww, _ = x.shape
        z = torch.empty(Bx, D, Hh, Ww, 48, dtype=F32, device=dev)
        self._conv("conv_out.conv.conv", x, causal, z, pre=pre)
        video = torch.empty(Bx, 3, D, 4 * Hh, 4 * Ww, dtype=F32, device=dev)
        ops.vae_unpatchify(z, video)
        return video.to(sample.dtype)

    def _timestep_embed_scaled(self, name: str, ts: Tensor, mult: float) -> Tensor:
        B = ts.numel()
        feat = torch.empty(B, 256, dtype=BF16, device=self.device)
        ops.timestep_embed(ts.reshape(-1), mult, 256, feat)
        l1, l2 = self._linears[name + ".timestep_embedder.linear_1"], self._linears[name + ".timestep_embedder.linear_2"]
        h = torch.empty(B, l1.weight.shape[0], dtype=BF16, device=self.device)
        ops.gemm(feat, l1.weight, l1.bias, h, _lib.EPI_SILU_BF16)
        e = torch.empty(B, l2.weight.shape[0], dtype=F32, device=self.device)
        ops.gemm(h, l2.weight, l2.bias, e, _lib.EPI_BIAS_F32)
        return e

    def decode_tiled(self, sample: Tensor, tiling_config: Optional[TilingConfig] = None, tiling_mode: str = "auto",
                     causal: bool = False, timestep: Optional[Tensor] = None, debug: bool = False,
                     on_frames_ready: Optional[Callable] = None) -> Tensor:
        """decoder.py:452-531: plain decode when the latents fit one tile, else decode_with_tiling."""
        if tiling_config is None:
            tiling_config = TilingConfig.default()
        _, _, f, h, w = sample.shape
        s_cfg, t_cfg = tiling_config.spatial_config, tiling_config.temporal_config
        spatial = s_cfg is not None and max(h, w) > s_cfg.tile_size_in_pixels // 32
        temporal = t_cfg is not None and f > t_cfg.tile_size_in_frames // 8
        if not spatial and not temporal:
            decoded = self(sample, causal=causal, timestep=timestep, debug=debug)
            if on_frames_ready is not None:
                on_frames_ready(decoded, 0)
            return decoded
        return decode_with_tiling(self, sample, tiling_config, 32, 8, causal=causal, timestep=timestep, on_frames_ready=on_frames_ready)


def load_vae_decoder(weights: Dict[str, Tensor], timestep_conditioning: Optional[bool] = None, device=None) -> LTX2VideoDecoder:
    """decoder.py:534-640 for tensors already in memory: keys under ``vae.decoder.`` / ``decoder.`` prefixes are accepted,
    ``vae.per_channel_statistics.{mean-of-means,std-of-means}`` map to ``latents_mean`` / ``latents_std``, PyTorch conv
    layouts are transposed; timestep conditioning follows from the presence of ``last_time_embedder`` weights."""
    out: Dict[str, Tensor] = {}
    for key, v in weights.items():
        k = key
        for pre in ("vae.decoder.", "decoder."):
            if k.startswith(pre):
                k = k[len(pre):]
        if k.endswith("per_channel_statistics.mean-of-means"):
            k = "latents_mean"
        elif k.endswith("per_channel_statistics.std-of-means"):
            k = "latents_std"
        out[k] = v
    if timestep_conditioning is None:
        timestep_conditioning = any(k.startswith("last_time_embedder") for k in out)
    model = LTX2VideoDecoder(timestep_conditioning=timestep_conditioning, device=device)
    params = model.parameters()
    model.load_weights({k: v for k, v in out.items() if k in params}, strict=True)
    return model
