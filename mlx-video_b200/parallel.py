"""Multi-GPU partitioning of the denoise step on one NVLink/NVSwitch box: one process per GPU,
``torch.distributed`` (NCCL) for the exchanges.

The reference is single-device: no collective exists to mirror, so this is new design.

* **Ulysses sequence parallelism** (``UlyssesGroup``).  Video tokens are sharded across P ranks; every
  row-wise operator (AdaLN, norms, projections, FFN, residuals, output head) runs on T/P local rows with
  replicated weights and needs no communication.  Only video self-attention mixes tokens:
      local QKV GEMM -> q/k RMSNorm over the FULL hidden row + RoPE on local rows (the norm statistic spans
      all heads, SURVEY F6, so it must precede the scatter) written straight into a head-grouped send buffer
      -> ONE all-to-all (Q|K|V packed): [T/P, H, dh] -> [T, H/P, dh]
      -> attention over all T tokens for H/P heads
      -> ONE all-to-all back -> the to_out GEMM reads the head-group-major result in place (3-D TMA map).
  Text cross-attention and audio->video attention have small replicated K/V: no communication.  The audio
  stream (Ta ~ 68 tokens) is replicated; video->audio attention all-gathers the projected video K/V.
* **CFG parallelism** (``CFGParallel``).  cond / uncond forwards of the dev pipeline are independent given the
  latents (generate.py:1258-1283): two rank groups run one each and swap the (B, T, 128) velocity.
"""
from __future__ import annotations

from dataclasses import replace
from typing import Dict, List, Optional, Tuple

import os

import torch
import torch.distributed as dist

from . import _lib
from .transformer import BF16, Modality, Workspace

Tensor = torch.Tensor


class CFGParallel:
    """Ranks [0, world/2) run the conditional forward, ranks [world/2, world) the unconditional one."""

    def __init__(self, world: int, rank: int) -> None:
        if world % 2:
            raise ValueError("CFG parallelism needs an even number of ranks")
        half = world // 2
        self.is_cond = rank < half
        self.partner = rank + half if self.is_cond else rank - half
        self.pair_group = None
        for r in range(half):  # every rank creates every group, in the same order
            g = dist.new_group([r, r + half])
            if r == min(rank, self.partner):
                self.pair_group = g

    def exchange(self, mine: Tensor) -> Tuple[Tensor, Tensor]:
        """-> (velocity_cond, velocity_uncond), both on every rank of the pair."""
        mine = mine.contiguous()
        both = [torch.empty_like(mine), torch.empty_like(mine)]
        dist.all_gather(both, mine, group=self.pair_group)
        return both[0], both[1]  # group ranks are ordered (cond, uncond)


class PeerMemory:
    """NVLink peer-mapped buffers of one rank group (torch symmetric memory supplies the allocation and the
    address exchange; every load / store / flag on them is issued by libltxb kernels)."""

    def __init__(self, group, size: int, index: int) -> None:
        import torch.distributed._symmetric_memory as symm

        self._symm, self.group, self.size, self.index = symm, group, size, index
        self._bufs: Dict[tuple, tuple] = {}
        self.epoch_counter = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", torch.cuda.current_device()))
        flags = symm.empty(64, dtype=torch.int32, device=torch.device("cuda", torch.cuda.current_device()))
        flags.zero_()
        torch.cuda.synchronize()
        handle = symm.rendezvous(flags, group)
        self._flags, self._flags_handle = flags, handle
        self.flag_ptrs = [int(p) for p in handle.buffer_ptrs]
        # the same barrier folded into the consumer kernel's prologue (ltxb_peer_sync): one struct, kept alive here
        self.done_counter = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", torch.cuda.current_device()))
        self.sync = _lib.PeerSync()
        for i, p in enumerate(self.flag_ptrs):
            self.sync.flags[i] = p
        self.sync.n_peers, self.sync.my_rank = size, index
        self.sync.epoch_counter, self.sync.done_counter = self.epoch_counter.data_ptr(), self.done_counter.data_ptr()
        dist.barrier(group=group)  # every rank has zeroed its flags before anyone raises one

    def buffer(self, tag: str, shape, dtype):
        """-> (local tensor, [base address of that buffer on every rank of the group]).  Collective on first use."""
        key = (tag, tuple(shape), dtype)
        hit = self._bufs.get(key)
        if hit is None:
            t = self._symm.empty(*shape, dtype=dtype, device=torch.device("cuda", torch.cuda.current_device()))
            handle = self._symm.rendezvous(t, self.group)
            hit = self._bufs[key] = (t, [int(p) for p in handle.buffer_ptrs], handle)
        return hit[0], hit[1]

    def barrier(self) -> None:
        from . import ops

        ops.peer_barrier(self.flag_ptrs, self.index, self.epoch_counter)

    def all_gather(self, tag: str, block: Tensor) -> Tensor:
        """Every rank's ``block`` -> (P, *block.shape) on every rank, as NVLink stores + one flag barrier.  The
        barrier BEFORE the stores keeps a fast rank from overwriting a buffer a slow rank is still reading."""
        from . import ops

        block = block.contiguous()
        out, ptrs = self.buffer(tag, (self.size, *block.shape), block.dtype)
        nbytes = block.numel() * block.element_size()
        self.barrier()
        ops.peer_broadcast(block, [p + self.index * nbytes for p in ptrs])
        self.barrier()
        return out


class UlyssesGroup:
    """Head-scatter / sequence-gather all-to-all around video self-attention over ``ranks``.

    ``fused=True`` (default on GPUs): the two all-to-alls of a block are not collective calls at all — the
    norm+RoPE kernel stores each head group straight into the destination rank's receive buffer, the attention
    kernel stores each output row straight into its owner's buffer, and a one-warp flag barrier on peer memory
    orders producers and consumers.  ``fused=False``: NCCL ``all_to_all_single`` between the same kernels."""

    def __init__(self, ranks: List[int], rank: int, group=None, fused: Optional[bool] = None) -> None:
        self.ranks, self.size = list(ranks), len(ranks)
        self.index = self.ranks.index(rank)
        self.group = group
        self._rope_cache: Dict[int, Tuple[Tensor, Tensor, Tuple[Tensor, Tensor]]] = {}
        self.pending_sync = None
        self.peers: Optional[PeerMemory] = None
        if fused is None:
            fused = torch.cuda.is_available() and dist.get_backend(group) == "nccl"
        if fused:
            self.peers = PeerMemory(dist.group.WORLD if group is None else group, self.size, self.index)

    # ------------------------------------------------------------------ token sharding (pure torch + dist)
    def local_slice(self, T: int) -> slice:
        if T % self.size:
            raise _lib.LtxbError(f"{T} video tokens do not split across {self.size} sequence-parallel ranks")
        tl = T // self.size
        return slice(self.index * tl, (self.index + 1) * tl)

    def shard_modality(self, m: Modality) -> Modality:
        T = m.latent.shape[1]
        sl = self.local_slice(T)
        ts = m.timesteps if m.timesteps.numel() == m.latent.shape[0] else m.timesteps[:, sl]
        pe = m.positional_embeddings
        if pe is not None:
            key = pe[0].data_ptr()
            hit = self._rope_cache.get(key)
            if hit is None or hit[0] is not pe[0] or hit[1] is not pe[1]:
                hit = (pe[0], pe[1], (pe[0][:, :, sl].contiguous(), pe[1][:, :, sl].contiguous()))
                # captured CUDA graphs hold the ADDRESSES of these slices: a slice must outlive every graph that was captured
                # with it, so entries are only dropped oldest-first beyond a handful of live tables (two stages x two
                # modalities in the reference pipelines)
                if len(self._rope_cache) >= 8:
                    self._rope_cache.pop(next(iter(self._rope_cache)))
                self._rope_cache[key] = hit
            pe = hit[2]
        return replace(m, latent=m.latent[:, sl].contiguous(), timesteps=ts.contiguous(), positions=m.positions[:, :, sl].contiguous(),
                       positional_embeddings=pe)

    def shard_inputs(self, video: Optional[Modality], audio: Optional[Modality]):
        if video is not None and video.latent.shape[0] != 1:
            raise _lib.LtxbError("sequence parallelism runs one video per group (use CFG parallelism or separate calls for B > 1)")
        return (None if video is None else self.shard_modality(video)), audio  # the short audio stream is replicated

    def gather_tokens(self, x: Tensor) -> Tensor:
        """(B, T/P, C) on every rank -> (B, T, C)."""
        x = x.contiguous()
        if self.peers is not None and x.shape[0] == 1 and (x.numel() * x.element_size()) % 16 == 0:
            return self.peers.all_gather("tokens", x[0]).reshape(1, self.size * x.shape[1], x.shape[2]).clone()
        parts = [torch.empty_like(x) for _ in range(self.size)]
        dist.all_gather(parts, x, group=self.group)
        return torch.cat(parts, dim=1)

    def gather_outputs(self, vx: Optional[Tensor], ax: Optional[Tensor]):
        return (None if vx is None else self.gather_tokens(vx)), ax

    # ------------------------------------------------------------------ the exchanges
    def all_to_all(self, send: Tensor, recv: Tensor) -> Tensor:
        dist.all_to_all_single(recv, send, group=self.group)
        return recv

    def self_attention(self, attn, ws: Workspace, tag: str, xq: Tensor, B: int, Tl: int, pe) -> Tuple[Tensor, int]:
        """Sequence-parallel attn1: xq bf16 [Tl, D] local rows -> (head-group-major attention output
        bf16 [P, Tl, (H/P)*dh], group width) for the to_out GEMM."""
        from . import ops

        P, H, dh, inner, dev = self.size, attn.heads, attn.dim_head, attn.inner_dim, xq.device
        if B != 1 or H % P:
            raise _lib.LtxbError(f"sequence parallelism needs B == 1 and heads ({H}) divisible by ranks ({P})")
        hp = H // P
        qkv = ws.get(tag + ".qkv", (Tl, 3 * inner), BF16, dev)
        ops.gemm(xq, attn.qkv_weight, attn.qkv_bias, qkv, const_w=True)
        if self.peers is not None:
            return self._self_attention_fused(attn, qkv, Tl, pe), hp * dh
        send = ws.get(tag + ".a2a_send", (P, Tl, 3, hp * dh), BF16, dev)
        cos, sin = (None, None) if pe is None else pe
        ops.qknorm_rope_scatter(qkv[:, :inner], send, 0, 3, P, 1, Tl, H, dh, attn.q_norm.weight, attn.q_norm.eps, cos, sin)
        ops.qknorm_rope_scatter(qkv[:, inner:2 * inner], send, 1, 3, P, 1, Tl, H, dh, attn.k_norm.weight, attn.k_norm.eps, cos, sin)
        ops.qknorm_rope_scatter(qkv[:, 2 * inner:], send, 2, 3, P, 1, Tl, H, dh, None, 0.0)
        recv = self.all_to_all(send, ws.get(tag + ".a2a_recv", (P, Tl, 3, hp * dh), BF16, dev))
        T = P * Tl
        full = recv.view(T, 3 * hp * dh)  # chunk i came from rank i = tokens [i*Tl, (i+1)*Tl): already in token order
        q, k, v = full[:, :hp * dh], full[:, hp * dh:2 * hp * dh], full[:, 2 * hp * dh:]
        o = attn.sdpa(ws, tag + ".sp", q, k, v, 1, T, T, None, heads=hp)  # [T, hp*dh]
        back = self.all_to_all(o.view(P, Tl, hp * dh), ws.get(tag + ".a2a_back", (P, Tl, hp * dh), BF16, dev))
        return back, hp * dh

    def _self_attention_fused(self, attn, qkv: Tensor, Tl: int, pe) -> Tensor:
        """The exchange as NVLink stores of the producing kernels (see class docstring).  Receive buffers are
        single: the two barriers of a block already order block b+1's remote writes after block b's reads."""
        import math

        from . import ops

        P, H, dh, inner, me = self.size, attn.heads, attn.dim_head, attn.inner_dim, self.index
        hp, w = H // P, (H // P) * dh
        T = P * Tl
        recv, recv_ptrs = self.peers.buffer("qkv", (P, Tl, 3, w), BF16)   # chunk i: rank i's tokens, my heads
        back, back_ptrs = self.peers.buffer("o", (P, Tl, w), BF16)         # chunk j: head group j, my tokens
        cos, sin = (None, None) if pe is None else pe
        chunk = Tl * 3 * w * 2  # bytes of one rank's chunk in a qkv receive buffer
        fused_qkv = attn.q_norm.weight.data_ptr() == attn.qk_norm_weight[0].data_ptr() and attn.k_norm.weight.data_ptr() == attn.qk_norm_weight[1].data_ptr()
        if fused_qkv:  # q | k | v in one launch (three column segments of the fused QKV buffer)
            bases = [recv_ptrs[j] + me * chunk for j in range(P)]
            ops.qkv_norm_rope_scatter_peers(qkv, inner, bases, Tl, H, dh, 3 * w, w, attn.qk_norm_weight, attn.q_norm.eps, cos, sin)
        else:
            for slot, (lo, norm) in enumerate(((0, attn.q_norm), (inner, attn.k_norm), (2 * inner, None))):
                bases = [recv_ptrs[j] + me * chunk + slot * w * 2 for j in range(P)]
                ops.qknorm_rope_scatter_peers(qkv[:, lo:lo + inner], bases, 1, Tl, H, dh, 3 * w, None if norm is None else norm.weight,
                                              0.0 if norm is None else norm.eps, None if norm is None else cos, None if norm is None else sin)
        # Two flag barriers per block: "every rank's q/k/v have landed here" before the attention reads them, "every rank's
        # attention rows for my tokens have landed here" before the out-projection reads those.  Each is a one-warp launch
        # of its own; LTXB_FOLD_BARRIERS=1 runs them in the prologue of the consumer kernel instead (ltxb_peer_sync: the
        # attention kernel, and the out-projection GEMM, which picks self.pending_sync up).  Measured on 2 GPUs, same box,
        # alternating runs: 23.66 / 23.78 ms per step folded against 23.41 / 23.60 with the launches — every block of the
        # consumer polls the flags and the last one to leave advances the epoch, which costs what the launch saved — so the
        # folded form stays opt-in (parity: scripts/sp_check.py passes either way).
        fold = os.environ.get("LTXB_FOLD_BARRIERS", "0") == "1"
        if not fold:
            self.peers.barrier()
        full = recv.view(T, 3 * w)
        o_bases = [back_ptrs[i] + me * Tl * w * 2 for i in range(P)]
        ops.attention_peers(full[:, :w], full[:, w:2 * w], full[:, 2 * w:], o_bases, Tl, w, T, T, hp, dh, 1.0 / math.sqrt(dh),
                            peer_sync=self.peers.sync if fold else None)
        if fold:
            self.pending_sync = self.peers.sync
        else:
            self.peers.barrier()
        return back

    def take_pending_sync(self):
        """The barrier the next reader of the exchange's result must run (None: already done)."""
        s, self.pending_sync = self.pending_sync, None
        return s

    def video_to_audio(self, attn, ws: Workspace, a_in: Tensor, v_in: Tensor, Ba: int, Ta: int, Tl: int, ax: Tensor, a, v,
                       gate: Tensor, gate_table: Tensor, row_div: int, row_index: Optional[Tensor]) -> None:
        """v2a (transformer.py:326-339) with the video rows sharded: queries = the replicated audio stream, keys / values =
        THIS rank's video rows.  Every rank attends over its own key slice and leaves the result un-normalised
        (``ltxb_attention_partial``: O~, stabiliser, row sum — 264 bytes per (head, audio row)); the P slices are
        all-gathered (0.57 MB per rank and block at 68 audio tokens, instead of 42 MB of projected video K/V at 5184 tokens)
        and merged by log-sum-exp (``ltxb_attention_merge``) on every rank, which keeps the replicated audio stream
        identical everywhere.  LTXB_V2A_LSE=0: the round-1 path (all-gather the projected K/V, attend redundantly)."""
        import math
        import os

        from . import ops

        inner, dev, H, dh = attn.inner_dim, a_in.device, attn.heads, attn.dim_head
        q, k, vv = attn.project(ws, "av.v2a", a_in, Ba, Ta, v_in, Tl, a.cross_positional_embeddings, v.cross_positional_embeddings)
        if os.environ.get("LTXB_V2A_LSE", "1") != "0" and Ta <= 256:
            n = ops.attention_partial_floats(Ba, Ta, H, dh)
            part = ws.get("av.v2a.part", (n,), torch.float32, dev)
            ops.attention_partial(q, k, vv, part, Ba, Ta, Tl, H, dh, 1.0 / math.sqrt(dh))
            if self.peers is not None:
                parts = self.peers.all_gather("av.v2a.parts", part)
            else:
                parts = ws.get("av.v2a.parts", (self.size, n), torch.float32, dev)
                dist.all_gather_into_tensor(parts, part, group=self.group)
            o = ws.get("av.v2a.o", (Ba * Ta, inner), BF16, dev)
            ops.attention_merge(parts.view(self.size, n), o, Ba, Ta, H, dh)
        else:
            kv_local = ws.get("av.v2a.kv", (Ba * Tl, 2 * inner), BF16, dev)  # the buffer k / vv are views of
            if self.peers is not None:
                parts = self.peers.all_gather("av.v2a.kv_all", kv_local)
            else:
                parts = ws.get("av.v2a.kv_all", (self.size, Ba * Tl, 2 * inner), BF16, dev)
                dist.all_gather_into_tensor(parts, kv_local, group=self.group)
            full = parts.view(self.size * Tl, 2 * inner)
            o = attn.sdpa(ws, "av.v2a", q, full[:, :inner], full[:, inner:], Ba, Ta, self.size * Tl, None)
        ops.gemm(o, attn.to_out.weight, attn.to_out.bias, ax, _lib.EPI_RESID_GATE_F32, resid=ax, gate=gate,
                 gate_table=gate_table, gate_row_div=row_div, gate_row_index=row_index, const_w=True)


class ParallelLayout:
    """How ``world`` ranks are used: optional CFG split (2 groups) x Ulysses group inside each."""

    def __init__(self, cfg: Optional[CFGParallel], ulysses: Optional[UlyssesGroup]) -> None:
        self.cfg, self.ulysses = cfg, ulysses

    def attach(self, model) -> None:
        model.seq_parallel = self.ulysses

    def describe(self) -> str:
        parts = []
        if self.cfg is not None:
            parts.append("cfg2")
        if self.ulysses is not None:
            parts.append(f"ulysses{self.ulysses.size}" + ("(nvlink-fused)" if self.ulysses.peers is not None else "(nccl)"))
        return "x".join(parts) or "single"


def plan(world: int, use_cfg: bool) -> Tuple[bool, int]:
    """(split cond/uncond across two rank groups?, Ulysses degree)."""
    if world < 1 or (world & (world - 1)):
        raise ValueError("world size must be a power of two (1, 2, 4, 8 GPUs of one box)")
    cfg = use_cfg and world >= 2
    return cfg, (world // 2 if cfg else world)


def make_layout(world: int, rank: int, use_cfg: bool, fused: Optional[bool] = None) -> ParallelLayout:
    cfg_split, sp = plan(world, use_cfg)
    cfg = CFGParallel(world, rank) if cfg_split else None
    uly = None
    if sp > 1:
        n_groups = world // sp
        for gi in range(n_groups):  # all ranks create all groups in the same order
            ranks = list(range(gi * sp, (gi + 1) * sp))
            g = dist.new_group(ranks) if n_groups > 1 else None
            if rank in ranks:
                uly = UlyssesGroup(ranks, rank, g, fused)
    return ParallelLayout(cfg, uly)
