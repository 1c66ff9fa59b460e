"""Quantised weights kept PACKED beside their bf16 expansion (row N3: ``nn.QuantizedLinear``, ltx.py:641-725).

``LTXModel.load_weights(..., keep_packed=True)`` registers, for every MLX affine-quantised linear of the checkpoint, the
uint32 words and the scales / biases exactly as stored, against the bf16 tensor they were expanded into.  ``ops.gemm`` then
sends a few-row product over such a weight (M <= ``max_rows``: a sequence-parallel shard, the audio stream, AdaLN rows — the
regime where the weight stream from HBM bounds the GEMM) to ``ltxb_gemm_qw_bf16``, which streams a quarter / half of the
bytes and is bit-identical to the bf16 path; larger M keeps using the expanded copy (tensor-bound, packed operands buy
nothing there).  Adjacent registrations (q | k | v slices of one fused matrix) are merged so the fused views resolve too.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch


@dataclass
class _Entry:
    base: int        # data_ptr of the bf16 expansion's first row
    pitch: int       # bytes between its rows
    rows: int
    cols: int
    packed: torch.Tensor   # int32 / uint32 [rows, cols * bits / 32]
    scales: torch.Tensor   # [rows, cols / group]
    biases: torch.Tensor
    group: int
    bits: int
    keep: torch.Tensor     # the bf16 tensor (keeps the address alive)


class PackedWeights:
    def __init__(self) -> None:
        self.entries: List[_Entry] = []
        self.max_rows = 256  # the packed kernel's limit (tensor-memory ring behind a 256-column accumulator)
        self._cache: Dict[tuple, Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor, int, int]]] = {}

    def __len__(self) -> int:
        return len(self.entries)

    def clear(self) -> None:
        self.entries.clear()
        self._cache.clear()

    def register(self, w: torch.Tensor, packed: torch.Tensor, scales: torch.Tensor, biases: torch.Tensor, group: int, bits: int) -> None:
        assert w.dim() == 2 and w.stride(1) == 1 and packed.shape[0] == w.shape[0] and bits in (2, 4, 8)
        if bits == 2:  # ltxb_gemm_qw_bf16 builds 4 and 8 bits; 2-bit weights stay on the expanded copy
            return
        self.entries.append(_Entry(w.data_ptr(), w.stride(0) * w.element_size(), w.shape[0], w.shape[1], packed.contiguous(),
                                   scales.contiguous(), biases.contiguous(), group, bits, w))
        self._cache.clear()

    def merge_adjacent(self) -> None:
        """Row-adjacent entries of one storage with the same layout (q | k | v) become one entry."""
        self.entries.sort(key=lambda e: e.base)
        merged: List[_Entry] = []
        for e in self.entries:
            m = merged[-1] if merged else None
            if (m is not None and m.base + m.rows * m.pitch == e.base and (m.pitch, m.cols, m.group, m.bits) == (e.pitch, e.cols, e.group, e.bits)
                    and m.scales.dtype == e.scales.dtype and m.packed.device == e.packed.device):
                merged[-1] = _Entry(m.base, m.pitch, m.rows + e.rows, m.cols, torch.cat([m.packed, e.packed]), torch.cat([m.scales, e.scales]),
                                    torch.cat([m.biases, e.biases]), m.group, m.bits, m.keep)
            else:
                merged.append(e)
        self.entries = merged
        self._cache.clear()

    def lookup(self, w: torch.Tensor):
        """-> (packed rows, scales rows, biases rows, group, bits) for a bf16 weight view that is a row range of a registered
        matrix, else None."""
        if not self.entries:
            return None
        key = (w.data_ptr(), w.shape[0], w.shape[1], w.stride(0))
        if key in self._cache:
            return self._cache[key]
        hit = None
        ptr, pitch = w.data_ptr(), w.stride(0) * w.element_size()
        for e in self.entries:
            off = ptr - e.base
            if off < 0 or off >= e.rows * e.pitch or pitch != e.pitch or w.shape[1] != e.cols or off % e.pitch:
                continue
            r0 = off // e.pitch
            if r0 + w.shape[0] <= e.rows and e.packed.device == w.device:
                hit = (e.packed[r0:r0 + w.shape[0]], e.scales[r0:r0 + w.shape[0]], e.biases[r0:r0 + w.shape[0]], e.group, e.bits)
            break
        self._cache[key] = hit
        return hit

    def nbytes(self) -> int:
        return sum(e.packed.numel() * 4 + 2 * e.scales.numel() * e.scales.element_size() for e in self.entries)


REGISTRY = PackedWeights()
