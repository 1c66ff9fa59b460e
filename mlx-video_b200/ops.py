"""Tensor-level wrappers over the C ABI: torch tensors in, raw device pointers + the current CUDA
stream out.  PyTorch is used for memory and streams only; all arithmetic happens in ``libltxb.so``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import Epilogue, check, lib
from .packed import REGISTRY as _packed

_device_set: set = set()


launches = 0          # kernels launched by the library since import (refreshed after every call; bench.py reports the delta)
_profile = None       # optional list: when set, every launch is bracketed by CUDA events -> (name, flops_or_bytes, start, end)


_profile_shapes: list = []  # (M, N, K) of every ltxb_gemm_bf16 launch recorded while profiling


def profile(enable: bool):
    """Per-launch CUDA-event timing on the launching stream (bench.py's roofline leg). Returns the record list."""
    global _profile, _profile_shapes
    _profile = [] if enable else None
    _profile_shapes = []
    return _profile


def _call(name: str, work: float, *args) -> None:
    """One C-ABI call == one kernel launch on torch's current stream."""
    global launches
    fn = getattr(lib, name)
    if _profile is None:
        rc = fn(*args)
        launches = lib.ltxb_kernel_launches()
        check(rc, name)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = fn(*args)
    e1.record()
    launches = lib.ltxb_kernel_launches()
    check(rc, name)
    _profile.append((name, work, e0, e1))
    if name == "ltxb_gemm_bf16":  # (M, N, K) of the launch, for the weight-stream roofline of few-row steps
        _profile_shapes.append(tuple(int(v) for v in args[6:9]))


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


_attn_workspaces: dict = {}  # device index -> key-split scratch of the attention kernel (ragged last wave)
_gemm_workspaces: dict = {}  # device index -> the stream-K scratch buffer registered with the library


def _prep(t: torch.Tensor) -> None:
    """Fail loudly on anything that is not a CUDA tensor; bind the library to the tensor's device."""
    if not t.is_cuda:
        raise _lib.LtxbError("ltxb ops need CUDA tensors; there is no CPU fallback on this path")
    idx = t.device.index
    if idx not in _device_set:
        check(lib.ltxb_set_device(idx), "ltxb_set_device")
        check(lib.ltxb_device_check(), "ltxb_device_check")
        _device_set.clear()
        _device_set.add(idx)
        if idx not in _gemm_workspaces:  # torch owns the memory, the library only borrows it
            ws = torch.empty(lib.ltxb_gemm_workspace_bytes(), dtype=torch.uint8, device=t.device)
            check(lib.ltxb_gemm_set_workspace(ws.data_ptr(), ws.numel(), _stream()), "ltxb_gemm_set_workspace")
            _gemm_workspaces[idx] = ws
        if idx not in _attn_workspaces:
            ws = torch.empty(lib.ltxb_attention_workspace_bytes(), dtype=torch.uint8, device=t.device)
            check(lib.ltxb_attention_set_workspace(ws.data_ptr(), ws.numel()), "ltxb_attention_set_workspace")
            _attn_workspaces[idx] = ws


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _rows(t: torch.Tensor) -> int:
    return t.numel() // t.shape[-1]


def _ld(t: torch.Tensor) -> int:
    """Leading dimension of a (possibly sliced) 2-D-viewable tensor whose last dim is contiguous."""
    assert t.stride(-1) == 1, "last dimension must be contiguous"
    if t.dim() == 1:
        return t.shape[0]
    ld = t.stride(-2)
    # all leading dims must collapse onto a single row stride
    expect = ld
    for d in range(t.dim() - 2, -1, -1):
        if t.shape[d] != 1:
            assert t.stride(d) == expect, f"tensor with strides {t.stride()} is not a strided 2-D view"
        expect *= t.shape[d]
    return ld


def _epilogue(mode, N, bias, resid, gate, gate_table, gate_row_div, gate_row_index, const_w, peer_sync=None) -> "Epilogue":
    epi = Epilogue()
    if peer_sync is not None:  # a _lib.PeerSync the caller keeps alive (PeerMemory.sync): flag barrier before the first read of A
        epi.peer_sync = C.pointer(peer_sync)
    epi.mode = mode
    epi.gate_row_div = gate_row_div
    epi.flags = 1 if const_w else 0
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
        epi.bias = bias.data_ptr()
    if mode == _lib.EPI_RESID_GATE_F32:
        assert resid is not None and resid.dtype == torch.float32
        epi.resid = resid.data_ptr()
        epi.ldr = _ld(resid)
        if gate is not None:
            assert gate.dtype == torch.float32 and gate.shape[-1] == N
            epi.gate = gate.data_ptr()
            epi.gate_ld = _ld(gate)
            epi.gate_row_index = _ptr(gate_row_index)
            if gate_table is not None:
                assert gate_table.dtype == torch.float32 and gate_table.numel() == N
                epi.gate_table = gate_table.data_ptr()
    return epi


def gemm(
    a: torch.Tensor,
    w: torch.Tensor,
    bias: Optional[torch.Tensor],
    out: torch.Tensor,
    mode: int = _lib.EPI_BIAS_BF16,
    resid: Optional[torch.Tensor] = None,
    gate: Optional[torch.Tensor] = None,
    gate_table: Optional[torch.Tensor] = None,
    gate_row_div: int = 1,
    gate_row_index: Optional[torch.Tensor] = None,
    block_n: int = 0,
    cta_pair: int = -1,
    a_group_cols: int = 0,
    const_w: bool = False,
    peer_sync=None,
) -> torch.Tensor:
    """out = epilogue(a @ w.T).  a: bf16 [..., K]; w: bf16 [N, K] (nn.Linear layout); see ltxb.h.
    a_group_cols = g > 0: ``a`` is a contiguous [K / g, M, g] tensor (head-group-major, as the Ulysses gather
    all-to-all delivers it) standing for the [M, K] operand.  const_w: ``w`` is a model weight no kernel ahead in the
    stream writes (LTXB_GEMM_CONST_W): the few-row kernel may start streaming it before its predecessor has finished."""
    _prep(a)
    if a_group_cols > 0:
        assert a.dim() == 3 and a.is_contiguous() and a.shape[2] == a_group_cols
        M, K = a.shape[1], a.shape[0] * a.shape[2]
    else:
        M, K = _rows(a), a.shape[-1]
    N = w.shape[0]
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and w.shape[1] == K
    assert _rows(out) == M and out.shape[-1] == N
    if mode in (_lib.EPI_BIAS_F32, _lib.EPI_RESID_GATE_F32):
        assert out.dtype == torch.float32
    else:
        assert out.dtype == torch.bfloat16
    if block_n == 0 and cta_pair < 0 and M <= _packed.max_rows and N >= 128 and K >= 256 and len(_packed):  # = where ltxb_gemm_bf16 picks the few-row kernel
        hit = _packed.lookup(w)  # a quantised linear kept packed (packed.py): stream the packed words instead
        if hit is not None:
            return gemm_qw(a, hit[0], hit[1], hit[2], hit[3], hit[4], bias, out, mode, resid, gate, gate_table, gate_row_div,
                           gate_row_index, 0, a_group_cols, const_w, peer_sync)
    epi = _epilogue(mode, N, bias, resid, gate, gate_table, gate_row_div, gate_row_index, const_w, peer_sync)
    lda = _ld(a)
    if a_group_cols > 0:
        epi.a_group_cols, epi.a_group_stride, lda = a_group_cols, M * a_group_cols, a_group_cols
    _call("ltxb_gemm_bf16", 2.0 * M * N * K, a.data_ptr(), lda, w.data_ptr(), _ld(w), out.data_ptr(), _ld(out), M, N, K,
                            C.byref(epi), block_n, cta_pair, _stream())
    return out


def gemm_qw(
    a: torch.Tensor,
    packed: torch.Tensor,
    scales: torch.Tensor,
    biases: torch.Tensor,
    group_size: int,
    bits: int,
    bias: Optional[torch.Tensor],
    out: torch.Tensor,
    mode: int = _lib.EPI_BIAS_BF16,
    resid: Optional[torch.Tensor] = None,
    gate: Optional[torch.Tensor] = None,
    gate_table: Optional[torch.Tensor] = None,
    gate_row_div: int = 1,
    gate_row_index: Optional[torch.Tensor] = None,
    splits: int = 0,
    a_group_cols: int = 0,
    const_w: bool = False,
    peer_sync=None,
) -> torch.Tensor:
    """out = epilogue(a @ dequant(packed, scales, biases).T) for FEW rows (M <= 256) with the MLX affine-quantised weight
    kept packed in HBM (``ltxb_gemm_qw_bf16``): uint32 / int32 [N, K*bits/32], scales / biases bf16 or f32 [N, K/group_size].
    Bit-identical to ``dequant_affine`` followed by ``gemm(..., cta_pair=4, block_n=splits)``."""
    _prep(a)
    if a_group_cols > 0:
        assert a.dim() == 3 and a.is_contiguous() and a.shape[2] == a_group_cols
        M, K = a.shape[1], a.shape[0] * a.shape[2]
    else:
        M, K = _rows(a), a.shape[-1]
    N = packed.shape[0]
    for t in (packed, scales, biases):
        if not t.is_cuda:
            raise _lib.LtxbError("ltxb ops need CUDA tensors; there is no CPU fallback on this path")
    assert a.dtype == torch.bfloat16 and packed.dtype in (torch.uint32, torch.int32) and packed.dim() == 2 and packed.stride(1) == 1
    assert scales.dtype == biases.dtype and scales.dtype in (torch.bfloat16, torch.float32)
    assert scales.shape == biases.shape and scales.stride() == biases.stride() and scales.stride(1) == 1
    if packed.shape[1] != K * bits // 32 or scales.shape != (N, K // group_size):
        raise ValueError(f"quantised tensors {tuple(packed.shape)} / {tuple(scales.shape)} do not describe a {N}x{K} weight at "
                         f"{bits} bits, group {group_size}")
    assert _rows(out) == M and out.shape[-1] == N
    assert out.dtype == (torch.float32 if mode in (_lib.EPI_BIAS_F32, _lib.EPI_RESID_GATE_F32) else torch.bfloat16)
    epi = _epilogue(mode, N, bias, resid, gate, gate_table, gate_row_div, gate_row_index, const_w, peer_sync)
    lda = _ld(a)
    if a_group_cols > 0:
        epi.a_group_cols, epi.a_group_stride, lda = a_group_cols, M * a_group_cols, a_group_cols
    _call("ltxb_gemm_qw_bf16", 2.0 * M * N * K, a.data_ptr(), lda, packed.data_ptr(), packed.stride(0), scales.data_ptr(), biases.data_ptr(),
          scales.stride(0), int(scales.dtype == torch.float32), group_size, bits, out.data_ptr(), _ld(out), M, N, K, C.byref(epi), splits,
          _stream())
    return out


def rmsnorm_modulate(
    x: torch.Tensor,
    out: torch.Tensor,
    eps: float,
    mod: Optional[torch.Tensor] = None,
    scale_off: int = 0,
    shift_off: int = 0,
    table_scale: Optional[torch.Tensor] = None,
    table_shift: Optional[torch.Tensor] = None,
    row_div: int = 1,
    row_index: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    _prep(x)
    assert x.dtype == torch.float32 and out.dtype == torch.bfloat16
    R, D = _rows(x), x.shape[-1]
    _call("ltxb_rmsnorm_modulate", 0.0, x.data_ptr(), _ld(x), out.data_ptr(), _ld(out), R, D, eps, _ptr(mod),
                                   0 if mod is None else _ld(mod), scale_off, shift_off, _ptr(table_scale),
                                   _ptr(table_shift), row_div, _ptr(row_index), _stream())
    return out


def residual_rmsnorm_modulate(
    x: torch.Tensor,
    y: torch.Tensor,
    out: torch.Tensor,
    eps: float,
    mod: Optional[torch.Tensor] = None,
    gate_off: int = -1,
    scale_off: int = -1,
    shift_off: int = -1,
    table_gate: Optional[torch.Tensor] = None,
    table_scale: Optional[torch.Tensor] = None,
    table_shift: Optional[torch.Tensor] = None,
    row_div: int = 1,
    row_index: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """x (f32, in place) += y (bf16) * gate; out (bf16) = rms_norm(x) * (1 + scale) + shift — one pass (ltxb.h K3r).
    Offsets index the columns of ``mod``; a negative offset means that role takes nothing from ``mod``."""
    _prep(x)
    assert x.dtype == torch.float32 and y.dtype == torch.bfloat16 and out.dtype == torch.bfloat16
    R, D = _rows(x), x.shape[-1]
    assert _rows(y) == R and y.shape[-1] == D
    _call("ltxb_residual_rmsnorm_modulate", 0.0, x.data_ptr(), _ld(x), y.data_ptr(), _ld(y), out.data_ptr(), _ld(out), R, D,
          eps, _ptr(mod), 0 if mod is None else _ld(mod), gate_off, scale_off, shift_off, _ptr(table_gate),
          _ptr(table_scale), _ptr(table_shift), row_div, _ptr(row_index), _stream())
    return out


def layernorm_modulate(
    x: torch.Tensor,
    out: torch.Tensor,
    eps: float,
    emb: Optional[torch.Tensor],
    table_scale: Optional[torch.Tensor],
    table_shift: Optional[torch.Tensor],
    row_div: int = 1,
    row_index: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    _prep(x)
    assert x.dtype == torch.float32 and out.dtype == torch.bfloat16
    R, D = _rows(x), x.shape[-1]
    _call("ltxb_layernorm_modulate", 0.0, x.data_ptr(), _ld(x), out.data_ptr(), _ld(out), R, D, eps, _ptr(emb),
                                     0 if emb is None else _ld(emb), _ptr(table_scale), _ptr(table_shift), row_div,
                                     _ptr(row_index), _stream())
    return out


def gate_residual(
    x: torch.Tensor,
    y: torch.Tensor,
    gate: Optional[torch.Tensor] = None,
    gate_off: int = 0,
    gate_table: Optional[torch.Tensor] = None,
    row_div: int = 1,
    row_index: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    _prep(x)
    assert x.dtype == torch.float32 and y.dtype == torch.bfloat16
    R, D = _rows(x), x.shape[-1]
    _call("ltxb_gate_residual", 0.0, x.data_ptr(), _ld(x), y.data_ptr(), _ld(y), R, D, _ptr(gate),
                                0 if gate is None else _ld(gate), gate_off, _ptr(gate_table), row_div, _ptr(row_index),
                                _stream())
    return x


def lora_merge(w: torch.Tensor, delta: torch.Tensor, strength: float) -> torch.Tensor:
    """w (bf16 [R, C], row-strided view allowed) <- bf16(w + bf16(delta * strength)) in place (ltxb.h N3)."""
    _prep(w)
    assert w.dtype == torch.bfloat16 and delta.dtype == torch.float32 and w.dim() == 2 and delta.shape == w.shape
    assert w.stride(1) == 1 and delta.stride(1) == 1
    _call("ltxb_lora_merge_bf16", 0.0, w.data_ptr(), w.stride(0), delta.data_ptr(), delta.stride(0), w.shape[0], w.shape[1],
          strength, _stream())
    return w


def im2col_cl(x: torch.Tensor, out: torch.Tensor, kd: int, kh: int, kw: int) -> torch.Tensor:
    """x f32 channels-last (N, D, H, W, C) -> out bf16 [N*D*H*W, kd*kh*kw*C]: the conv's GEMM operand (ltxb.h N4)."""
    _prep(x)
    N, D, H, W, Cc = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous() and out.dtype == torch.bfloat16 and out.is_contiguous()
    assert out.shape == (N * D * H * W, kd * kh * kw * Cc)
    _call("ltxb_im2col_cl", 0.0, x.data_ptr(), out.data_ptr(), N, D, H, W, Cc, kd, kh, kw, _stream())
    return out


_gn_workspaces: dict = {}  # device index -> GroupNorm partial-sum scratch (grown on demand; torch owns the memory)


def groupnorm_silu(x: torch.Tensor, out: torch.Tensor, groups: int, eps: float, weight: torch.Tensor, bias: torch.Tensor,
                   resid: Optional[torch.Tensor] = None, silu: bool = True) -> torch.Tensor:
    """out = silu?(GroupNorm(x) * weight + bias [+ resid]) on f32 (N, S, C) (channels-last, any number of middle axes)."""
    _prep(x)
    for t in (x, out, weight, bias, resid):
        assert t is None or (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous())
    N, Cc = x.shape[0], x.shape[-1]
    S = x.numel() // (N * Cc)
    need = lib.ltxb_groupnorm_workspace_bytes(N, S, groups)
    ws = _gn_workspaces.get(x.device.index)
    if ws is None or ws.numel() < need:
        ws = _gn_workspaces[x.device.index] = torch.empty(max(need, 1 << 16), dtype=torch.uint8, device=x.device)
    _call("ltxb_groupnorm_silu", 0.0, x.data_ptr(), out.data_ptr(), N, S, Cc, groups, eps, weight.data_ptr(), bias.data_ptr(),
          _ptr(resid), int(silu), ws.data_ptr(), ws.numel(), _stream())
    return out


def pixel_shuffle2(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """x f32 (F, H, W, 4*Co) -> out f32 (F, 2H, 2W, Co)."""
    _prep(x)
    F_, H, W, C4 = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous() and out.is_contiguous() and out.shape == (F_, 2 * H, 2 * W, C4 // 4)
    _call("ltxb_pixel_shuffle2", 0.0, x.data_ptr(), out.data_ptr(), F_, H, W, C4 // 4, _stream())
    return out


def latent_layout(x: torch.Tensor, out: torch.Tensor, scale: Optional[torch.Tensor], shift: Optional[torch.Tensor],
                  to_channels_last: bool) -> torch.Tensor:
    """(B, C, S) <-> (B, S, C) on f32, with the per-channel (un)normalisation of upsample_latents folded in."""
    _prep(x)
    for t in (x, out, scale, shift):
        assert t is None or (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous())
    if to_channels_last:
        B, Cc, S = x.shape
    else:
        B, S, Cc = x.shape
    _call("ltxb_latent_layout", 0.0, x.data_ptr(), out.data_ptr(), _ptr(scale), _ptr(shift), B, Cc, S, int(to_channels_last), _stream())
    return out


def dequant_affine(packed: torch.Tensor, scales: torch.Tensor, biases: torch.Tensor, out: torch.Tensor, group_size: int,
                   bits: int) -> torch.Tensor:
    """out (bf16 [R, C], row-strided view allowed) <- scales * q + biases of an MLX affine-quantised weight
    (packed uint32 / int32 [R, C*bits/32]; scales, biases bf16 or f32 [R, C/group_size]) (ltxb.h N3)."""
    _prep(out)
    for t in (packed, scales, biases):
        if not t.is_cuda:
            raise _lib.LtxbError("ltxb ops need CUDA tensors; there is no CPU fallback on this path")
    assert out.dtype == torch.bfloat16 and out.dim() == 2 and out.stride(1) == 1
    assert packed.dtype in (torch.uint32, torch.int32) and packed.dim() == 2 and packed.stride(1) == 1
    assert scales.dtype == biases.dtype and scales.dtype in (torch.bfloat16, torch.float32)
    assert scales.shape == biases.shape and scales.stride() == biases.stride() and scales.stride(1) == 1
    R, Cc = out.shape
    if packed.shape != (R, Cc * bits // 32) or scales.shape != (R, Cc // group_size):
        raise ValueError(f"quantised tensors {tuple(packed.shape)} / {tuple(scales.shape)} do not describe a "
                         f"{R}x{Cc} weight at {bits} bits, group {group_size}")
    _call("ltxb_dequant_affine_bf16", 0.0, packed.data_ptr(), packed.stride(0), scales.data_ptr(), biases.data_ptr(),
          scales.stride(0), int(scales.dtype == torch.float32), out.data_ptr(), out.stride(0), R, Cc, group_size, bits, _stream())
    return out


def vae_gather_rows(x: torch.Tensor, out: torch.Tensor, causal: bool, m0: int, rows: int, pre_op: bool = False,
                    table_scale: Optional[torch.Tensor] = None, table_shift: Optional[torch.Tensor] = None,
                    emb_scale: Optional[torch.Tensor] = None, emb_shift: Optional[torch.Tensor] = None, emb_ld: int = 0,
                    eps: float = 1e-8) -> torch.Tensor:
    """Rows [m0, m0+rows) of the CausalConv3d operand of x f32 channels-last (N, D, H, W, C) -> out bf16 [rows, 27*C] (ltxb.h N4)."""
    _prep(x)
    N, D, H, W, Cc = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous() and out.dtype == torch.bfloat16 and out.is_contiguous()
    assert out.shape[0] >= rows and out.shape[1] == 27 * Cc
    for t in (table_scale, table_shift, emb_scale, emb_shift):
        assert t is None or (t.is_cuda and t.dtype == torch.float32 and t.stride(-1) == 1)
    _call("ltxb_vae_gather_rows", 0.0, x.data_ptr(), out.data_ptr(), N, D, H, W, Cc, int(causal), m0, rows, _ptr(table_scale),
          _ptr(table_shift), _ptr(emb_scale), _ptr(emb_shift), emb_ld, eps, int(pre_op), _stream())
    return out


def vae_depth_to_space(y: torch.Tensor, x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """y f32 (N, D, H, W, 4C) conv output + x (N, D, H, W, C) conv input -> out (N, 2D-1, 2H, 2W, C/2) (ltxb.h N4)."""
    _prep(y)
    N, D, H, W, Cc = x.shape
    assert y.shape == (N, D, H, W, 4 * Cc) and out.shape == (N, 2 * D - 1, 2 * H, 2 * W, Cc // 2)
    for t in (y, x, out):
        assert t.dtype == torch.float32 and t.is_contiguous()
    _call("ltxb_vae_depth_to_space", 0.0, y.data_ptr(), x.data_ptr(), out.data_ptr(), N, D, H, W, Cc, _stream())
    return out


def vae_prepare_latent(sample: torch.Tensor, noise: Optional[torch.Tensor], noise_scale: float, std: torch.Tensor, mean: torch.Tensor,
                       out: torch.Tensor) -> torch.Tensor:
    """sample f32 channels-first (N, C, F, H, W) [+ noise] -> out channels-last (N, F, H, W, C), de-normalised (ltxb.h N4)."""
    _prep(sample)
    N, Cc = sample.shape[:2]
    S = sample.numel() // (N * Cc)
    for t in (sample, noise, std, mean, out):
        assert t is None or (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous())
    assert out.numel() == sample.numel() and out.shape[-1] == Cc
    _call("ltxb_vae_prepare_latent", 0.0, sample.data_ptr(), _ptr(noise), noise_scale, std.data_ptr(), mean.data_ptr(), out.data_ptr(),
          N, Cc, S, _stream())
    return out


def vae_unpatchify(z: torch.Tensor, video: torch.Tensor) -> torch.Tensor:
    """z f32 channels-last (N, F, H, W, 48) -> video channels-first (N, 3, F, 4H, 4W)."""
    _prep(z)
    N, F_, H, W, Cc = z.shape
    assert Cc == 48 and video.shape == (N, 3, F_, 4 * H, 4 * W) and z.is_contiguous() and video.is_contiguous()
    assert z.dtype == torch.float32 and video.dtype == torch.float32
    _call("ltxb_vae_unpatchify", 0.0, z.data_ptr(), video.data_ptr(), N, F_, H, W, _stream())
    return video


def vae_blend_tile(tile: torch.Tensor, at: int, ah: int, aw: int, mt: torch.Tensor, mh: torch.Tensor, mw: torch.Tensor,
                   output: torch.Tensor, weights: torch.Tensor, t0: int, h0: int, w0: int) -> None:
    """output (N,3,F,H,W) += tile[..., :at, :ah, :aw] * mt x mh x mw at (t0, h0, w0); weights (N,1,F,H,W) += mask."""
    _prep(tile)
    N, _, Ft, Ht, Wt = tile.shape
    _, _, F_, H, W = output.shape
    for t in (tile, mt, mh, mw, output, weights):
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
    assert mt.numel() >= at and mh.numel() >= ah and mw.numel() >= aw and weights.shape == (N, 1, F_, H, W)
    _call("ltxb_vae_blend_tile", 0.0, tile.data_ptr(), N, Ft, Ht, Wt, at, ah, aw, mt.data_ptr(), mh.data_ptr(), mw.data_ptr(),
          output.data_ptr(), weights.data_ptr(), F_, H, W, t0, h0, w0, _stream())


def vae_blend_normalize(output: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    _prep(output)
    N = output.shape[0]
    plane = output.numel() // (N * 3)
    assert output.dtype == torch.float32 and weights.dtype == torch.float32 and weights.numel() == N * plane
    _call("ltxb_vae_blend_normalize", 0.0, output.data_ptr(), weights.data_ptr(), N, plane, _stream())
    return output


def qknorm_rope(
    x: torch.Tensor,
    B: int,
    T: int,
    H: int,
    dh: int,
    weight: torch.Tensor,
    eps: float,
    cos: Optional[torch.Tensor] = None,
    sin: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """In place on a bf16 [B*T, H*dh] (strided) view."""
    _prep(x)
    assert x.dtype == torch.bfloat16 and weight.dtype == torch.float32
    b_pe = 1
    if cos is not None:
        assert cos.dtype == torch.float32 and sin.dtype == torch.float32 and cos.is_contiguous() and sin.is_contiguous()
        assert cos.shape[1:] == (H, T, dh // 2), f"rope table {tuple(cos.shape)} vs (B,{H},{T},{dh // 2})"
        b_pe = cos.shape[0]
    _call("ltxb_qknorm_rope", 0.0, x.data_ptr(), _ld(x), B, T, H, dh, weight.data_ptr(), eps, _ptr(cos), _ptr(sin), b_pe,
                              _stream())
    return x


def timestep_embed(t: torch.Tensor, scale: float, dim: int, out: torch.Tensor) -> torch.Tensor:
    _prep(t)
    assert t.dtype == torch.float32 and t.is_contiguous() and out.dtype == torch.bfloat16
    _call("ltxb_timestep_embed", 0.0, t.data_ptr(), t.numel(), scale, dim, out.data_ptr(), _ld(out), _stream())
    return out


def rope_table(positions: torch.Tensor, max_pos, freq: torch.Tensor, dim: int, H: int, use_middle: bool):
    _prep(positions)
    assert positions.dtype == torch.float32 and positions.is_contiguous() and positions.dim() == 4
    B, n_axes, T, two = positions.shape
    assert two == 2 and len(max_pos) == n_axes
    half = dim // 2
    cos = torch.empty((B, H, T, half // H), dtype=torch.float32, device=positions.device)
    sin = torch.empty_like(cos)
    mp = (C.c_float * n_axes)(*[float(v) for v in max_pos])
    _call("ltxb_rope_table", 0.0, positions.data_ptr(), B, n_axes, T, mp, freq.data_ptr(), freq.numel(), dim, H,
                             1 if use_middle else 0, cos.data_ptr(), sin.data_ptr(), _stream())
    return cos, sin


def silu_bf16(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    _prep(x)
    assert x.dtype == torch.bfloat16 and out.dtype == torch.bfloat16 and x.is_contiguous() and out.is_contiguous()
    _call("ltxb_silu_bf16", 0.0, x.data_ptr(), out.data_ptr(), x.numel(), _stream())
    return out


def cast_f32_to_bf16(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    _prep(x)
    assert x.dtype == torch.float32 and out.dtype == torch.bfloat16 and x.is_contiguous() and out.is_contiguous()
    _call("ltxb_cast_f32_to_bf16", 0.0, x.data_ptr(), out.data_ptr(), x.numel(), _stream())
    return out


def cast_bf16_to_f32(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    _prep(x)
    assert x.dtype == torch.bfloat16 and out.dtype == torch.float32 and x.is_contiguous() and out.is_contiguous()
    _call("ltxb_cast_bf16_to_f32", 0.0, x.data_ptr(), out.data_ptr(), x.numel(), _stream())
    return out


def attention(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    out: torch.Tensor,
    B: int,
    Tq: int,
    Tk: int,
    H: int,
    dh: int,
    scale: float,
    kv_bias: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """q: bf16 [B*Tq, H*dh] view, k/v: bf16 [B*Tk, H*dh] views (row-strided), out: bf16 [B*Tq, H*dh]."""
    _prep(q)
    assert q.dtype == k.dtype == v.dtype == out.dtype == torch.bfloat16
    if kv_bias is not None:
        assert kv_bias.dtype == torch.float32 and kv_bias.is_contiguous() and kv_bias.shape == (B, Tk)
    _call("ltxb_attention_fwd", 4.0 * B * H * Tq * Tk * dh, q.data_ptr(), _ld(q), k.data_ptr(), _ld(k), v.data_ptr(), _ld(v), out.data_ptr(),
                                _ld(out), B, Tq, Tk, H, dh, scale, _ptr(kv_bias), _stream())
    return out


def attention_partial_floats(B: int, Tq: int, H: int, dh: int) -> int:
    return int(lib.ltxb_attention_partial_floats(B, Tq, H, dh))


def attention_partial(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, part: torch.Tensor, B: int, Tq: int, Tk: int, H: int, dh: int,
                      scale: float) -> torch.Tensor:
    """Un-normalised attention over the given keys: part (f32, attention_partial_floats elements) <- [B*H*Tq][dh] | [B*H*Tq][2]."""
    _prep(q)
    assert q.dtype == k.dtype == v.dtype == torch.bfloat16 and part.dtype == torch.float32 and part.is_contiguous()
    assert part.numel() >= attention_partial_floats(B, Tq, H, dh)
    _call("ltxb_attention_partial", 4.0 * B * H * Tq * Tk * dh, q.data_ptr(), _ld(q), k.data_ptr(), _ld(k), v.data_ptr(), _ld(v), part.data_ptr(),
          B, Tq, Tk, H, dh, scale, _stream())
    return part


def attention_merge(parts: torch.Tensor, out: torch.Tensor, B: int, Tq: int, H: int, dh: int) -> torch.Tensor:
    """parts f32 [n_parts, block] (contiguous; block >= attention_partial_floats) -> out bf16 [B*Tq, H*dh] (log-sum-exp merge)."""
    _prep(parts)
    assert parts.dtype == torch.float32 and parts.dim() == 2 and parts.is_contiguous() and out.dtype == torch.bfloat16
    _call("ltxb_attention_merge", 0.0, parts.data_ptr(), parts.shape[1], parts.shape[0], out.data_ptr(), _ld(out), B, Tq, H, dh, _stream())
    return out


def euler_step(
    x: torch.Tensor,
    v_pos: torch.Tensor,
    sigma: float,
    sigma_next: float,
    v_neg: Optional[torch.Tensor] = None,
    cfg_scale: float = 1.0,
    sigma_tok: Optional[torch.Tensor] = None,
    mask: Optional[torch.Tensor] = None,
    clean: Optional[torch.Tensor] = None,
    x0_out: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """In-place sampler step on f32 [n_tok, C] latents (CFG combine + to_denoised + mask blend + Euler)."""
    _prep(x)
    for t in (x, v_pos, v_neg, sigma_tok, mask, clean, x0_out):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous())
    n_tok, Cc = _rows(x), x.shape[-1]
    _call("ltxb_euler_step", 0.0, x.data_ptr(), v_pos.data_ptr(), _ptr(v_neg), cfg_scale, _ptr(sigma_tok), sigma, sigma_next,
                             _ptr(mask), _ptr(clean), n_tok, Cc, _ptr(x0_out), _stream())
    return x


def timestep_groups(t: torch.Tensor, cap: int):
    """Device-side dedupe of per-token timesteps: returns (values f32 [cap], index i32 [n], count i32 [1])."""
    _prep(t)
    assert t.dtype == torch.float32 and t.is_contiguous()
    n = t.numel()
    values = torch.empty(cap, dtype=torch.float32, device=t.device)
    index = torch.empty(n, dtype=torch.int32, device=t.device)
    count = torch.empty(1, dtype=torch.int32, device=t.device)
    _call("ltxb_timestep_groups", 0.0, t.data_ptr(), n, cap, values.data_ptr(), index.data_ptr(), count.data_ptr(), _stream())
    return values, index, count


def qknorm_rope_segments(x: torch.Tensor, n_seg: int, seg_stride: int, B: int, T: int, H: int, dh: int, weight: torch.Tensor,
                         eps: float, cos: Optional[torch.Tensor] = None, sin: Optional[torch.Tensor] = None,
                         weight2: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ONE launch over ``n_seg`` column slices x[:, s*seg_stride : s*seg_stride + H*dh] of a bf16 buffer, slice s
    normalised with ``weight[s]`` (f32 [n_seg, H*dh], optionally times ``weight2[s]``) and rotated in place."""
    _prep(x)
    assert x.dtype == torch.bfloat16 and weight.dtype == torch.float32 and weight.is_contiguous()
    assert tuple(weight.shape) == (n_seg, H * dh), f"segment weights {tuple(weight.shape)} vs ({n_seg}, {H * dh})"
    assert (n_seg - 1) * seg_stride + H * dh <= x.shape[-1]
    if weight2 is not None:
        assert weight2.dtype == torch.float32 and weight2.is_contiguous() and weight2.shape == weight.shape
    b_pe = 1
    if cos is not None:
        assert cos.dtype == torch.float32 and sin.dtype == torch.float32 and cos.is_contiguous() and sin.is_contiguous()
        assert cos.shape[1:] == (H, T, dh // 2), f"rope table {tuple(cos.shape)} vs (B,{H},{T},{dh // 2})"
        b_pe = cos.shape[0]
    _call("ltxb_qknorm_rope_segments", 0.0, x.data_ptr(), _ld(x), n_seg, seg_stride, B, T, H, dh, weight.data_ptr(), H * dh,
          _ptr(weight2), eps, _ptr(cos), _ptr(sin), b_pe, _stream())
    return x


def qknorm_rope_scatter(x: torch.Tensor, out: torch.Tensor, slot: int, slots: int, groups: int, B: int, T: int, H: int,
                        dh: int, weight: Optional[torch.Tensor], eps: float, cos: Optional[torch.Tensor] = None,
                        sin: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q/k RMSNorm (+ split RoPE) of x bf16 [B*T, H*dh] written into slot ``slot`` of the head-grouped send
    buffer ``out`` bf16 [groups, B*T, slots, (H/groups)*dh]; weight None = plain copy (V)."""
    _prep(x)
    hp = H // groups
    assert x.dtype == torch.bfloat16 and out.dtype == torch.bfloat16 and out.is_contiguous()
    assert tuple(out.shape) == (groups, B * T, slots, hp * dh), f"send buffer {tuple(out.shape)}"
    b_pe = 1
    if cos is not None:
        assert cos.is_contiguous() and sin.is_contiguous() and cos.shape[1:] == (H, T, dh // 2)
        b_pe = cos.shape[0]
    base = out.data_ptr() + slot * hp * dh * 2
    _call("ltxb_qknorm_rope_scatter", 0.0, x.data_ptr(), _ld(x), base, slots * hp * dh, hp, B * T * slots * hp * dh, B, T, H, dh,
          _ptr(weight), eps, _ptr(cos), _ptr(sin), b_pe, _stream())
    return out


def _ptr_array(ptrs):
    return (C.c_void_p * len(ptrs))(*[int(p) for p in ptrs])


def qknorm_rope_scatter_peers(x: torch.Tensor, group_bases, B: int, T: int, H: int, dh: int, ldo: int,
                              weight: Optional[torch.Tensor], eps: float, cos: Optional[torch.Tensor] = None,
                              sin: Optional[torch.Tensor] = None) -> None:
    """q/k RMSNorm (+ RoPE) of x bf16 [B*T, H*dh]; head group g is stored to the raw device address
    ``group_bases[g]`` (a peer GPU's receive buffer, rows ``ldo`` elements apart).  weight None = copy (V)."""
    _prep(x)
    assert x.dtype == torch.bfloat16
    b_pe = 1
    if cos is not None:
        assert cos.is_contiguous() and sin.is_contiguous() and cos.shape[1:] == (H, T, dh // 2)
        b_pe = cos.shape[0]
    _call("ltxb_qknorm_rope_scatter_peers", 0.0, x.data_ptr(), _ld(x), _ptr_array(group_bases), len(group_bases), ldo, B, T, H, dh,
          _ptr(weight), eps, _ptr(cos), _ptr(sin), b_pe, _stream())


def qkv_norm_rope_scatter_peers(qkv: torch.Tensor, inner: int, group_bases, T: int, H: int, dh: int, ldo: int, slot_stride: int,
                                qk_weight: torch.Tensor, eps: float, cos: Optional[torch.Tensor] = None,
                                sin: Optional[torch.Tensor] = None) -> None:
    """q | k | v column segments of the fused QKV buffer (bf16 [T, 3*inner]) in one launch: q, k normalised with
    ``qk_weight`` (f32 [2, inner]) and rotated, v copied; head group g of segment s goes to ``group_bases[g] + s*slot_stride``."""
    _prep(qkv)
    assert qkv.dtype == torch.bfloat16 and qkv.shape[-1] == 3 * inner and inner == H * dh
    assert qk_weight.dtype == torch.float32 and qk_weight.is_contiguous() and tuple(qk_weight.shape) == (2, inner)
    b_pe = 1
    if cos is not None:
        assert cos.is_contiguous() and sin.is_contiguous() and cos.shape[1:] == (H, T, dh // 2)
        b_pe = cos.shape[0]
    _call("ltxb_qkv_norm_rope_scatter_peers", 0.0, qkv.data_ptr(), _ld(qkv), inner, _ptr_array(group_bases), len(group_bases), ldo, slot_stride,
          T, H, dh, qk_weight.data_ptr(), eps, _ptr(cos), _ptr(sin), b_pe, _stream())


def attention_peers(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, o_bases, rows_per_peer: int, ldo: int, Tq: int, Tk: int,
                    H: int, dh: int, scale: float, peer_sync=None) -> None:
    """Attention (B = 1) whose output rows [i*rows_per_peer, (i+1)*rows_per_peer) go to ``o_bases[i]``.  peer_sync (a
    _lib.PeerSync kept alive by the caller): q / k / v were stored by the peers — the flag barrier that orders those stores
    before this kernel's reads runs in its prologue instead of as a launch of its own."""
    _prep(q)
    assert q.dtype == k.dtype == v.dtype == torch.bfloat16
    if peer_sync is not None:
        _call("ltxb_attention_fwd_peers_sync", 4.0 * H * Tq * Tk * dh, q.data_ptr(), _ld(q), k.data_ptr(), _ld(k), v.data_ptr(), _ld(v),
              _ptr_array(o_bases), len(o_bases), rows_per_peer, ldo, Tq, Tk, H, dh, scale, C.pointer(peer_sync), _stream())
        return
    _call("ltxb_attention_fwd_peers", 4.0 * H * Tq * Tk * dh, q.data_ptr(), _ld(q), k.data_ptr(), _ld(k), v.data_ptr(), _ld(v),
          _ptr_array(o_bases), len(o_bases), rows_per_peer, ldo, Tq, Tk, H, dh, scale, _stream())


def peer_barrier(flag_ptrs, my_rank: int, epoch_counter: torch.Tensor) -> None:
    """Stream-ordered cross-GPU barrier on NVLink-mapped flag arrays (see ltxb.h)."""
    _prep(epoch_counter)
    assert epoch_counter.dtype == torch.int32
    _call("ltxb_peer_barrier", 0.0, _ptr_array(flag_ptrs), len(flag_ptrs), my_rank, epoch_counter.data_ptr(), _stream())


def peer_broadcast(src: torch.Tensor, dst_ptrs) -> None:
    """This rank's contiguous block -> the raw device addresses ``dst_ptrs`` (one per peer)."""
    _prep(src)
    assert src.is_contiguous()
    _call("ltxb_peer_broadcast", 0.0, src.data_ptr(), src.numel() * src.element_size(), _ptr_array(dst_ptrs), len(dst_ptrs), _stream())
