"""RoPE frequency tables for the LTX-2 DiT, built on the device by ``ltxb_rope_table``.

Mirrors ``precompute_freqs_cis`` of the reference (mlx_video/models/ltx/rope.py:364-416 dispatching to
:419-529).  The per-axis base frequencies are produced on the host exactly the way the reference does
(fp32 ``theta ** linspace(0, 1, N) * pi/2``, rope.py:446-457); position -> angle -> cos/sin and the
(B, H, T, dim/2H) SPLIT layout with its left identity padding (rope.py:499-516) are the kernel's.
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import torch

from . import ops
from ._lib import LtxbError
from .config import LTXRopeType


def rope_base_frequencies(theta: float, n_pos_dims: int, dim: int) -> torch.Tensor:
    """fp32 [dim // (2 n_pos_dims)] base frequencies (rope.py:446-457)."""
    n = max(dim // (2 * n_pos_dims), 1)
    ramp = torch.linspace(0.0, 1.0, n, dtype=torch.float32)
    return torch.pow(torch.tensor(theta, dtype=torch.float32), ramp) * (math.pi / 2)


def precompute_freqs_cis(
    indices_grid: torch.Tensor,
    dim: int,
    theta: float = 10000.0,
    max_pos: Optional[List[int]] = None,
    use_middle_indices_grid: bool = False,
    num_attention_heads: int = 32,
    rope_type: LTXRopeType = LTXRopeType.INTERLEAVED,
    double_precision: bool = False,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """positions (B, n_axes, T, 2) [start, end) bounds (or (B, n_axes, T)) -> cos, sin fp32 (B, H, T, dim/2H).

    ``double_precision`` only selects between two fp32 formulations in the reference (rope.py:428-431);
    both give the same table, so it is accepted and ignored.  Only the SPLIT layout — the one every
    pipeline of the reference configures (generate.py:2875) — has kernels; INTERLEAVED raises.
    """
    if max_pos is None:
        max_pos = [20, 2048, 2048]
    if rope_type != LTXRopeType.SPLIT:
        if not isinstance(rope_type, LTXRopeType):
            raise ValueError(f"Invalid rope type: {rope_type}")  # rope.py:30
        raise LtxbError(f"rope_type {rope_type} has no sm_100a kernel on this path; LTX-2 pipelines use SPLIT")
    grid = indices_grid
    if not grid.is_cuda:
        raise LtxbError("precompute_freqs_cis needs the position grid on the CUDA device (no CPU fallback)")
    if grid.dim() == 3:  # already collapsed to one coordinate per token
        grid = torch.stack([grid, grid], dim=-1)
        use_middle_indices_grid = False
    assert grid.dim() == 4 and grid.shape[-1] == 2, "position grid must be (B, n_axes, T, 2)"
    n_axes = grid.shape[1]
    assert n_axes == len(max_pos), "Number of position dimensions must match max_pos length"  # rope.py:228
    grid = grid.to(torch.float32).contiguous()
    key = (float(theta), n_axes, dim, str(grid.device))
    freq = _freq_cache.get(key)
    if freq is None:  # host-built once per geometry, then resident (no H2D copy in the steady state / under graph capture)
        freq = _freq_cache[key] = rope_base_frequencies(theta, n_axes, dim).to(grid.device)
    return ops.rope_table(grid, max_pos, freq, dim, num_attention_heads, use_middle_indices_grid)


_freq_cache: dict = {}
