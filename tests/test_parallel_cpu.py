"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU path (token sharding, the packed
head-scatter / sequence-gather all-to-all layout, CFG velocity exchange).  The CUDA kernels are replaced by
torch stand-ins that honour the SAME buffer layouts, so what is tested is the index math and the collective
sequence of mlx_video_b200/parallel.py — not the arithmetic (tests/test_gpu_* cover that)."""
import math
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import mlx_video_b200 as M
from mlx_video_b200 import _lib, ops, parallel
from mlx_video_b200.transformer import Workspace


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


# ---- torch stand-ins with the kernels' exact buffer contracts ------------------------------------------
def fake_gemm(a, w, bias, out, mode=_lib.EPI_BIAS_BF16, a_group_cols=0, **kw):
    if a_group_cols:
        a = a.permute(1, 0, 2).reshape(a.shape[1], -1)  # [K/g, M, g] -> [M, K]
    y = a.float() @ w.float().T + (0 if bias is None else bias)
    out.copy_(y.to(out.dtype))
    return out


def fake_scatter(x, out, slot, slots, groups, B, T, H, dh, weight, eps, cos=None, sin=None):
    y = x.float()
    if weight is not None:
        y = y * torch.rsqrt(y.pow(2).mean(-1, keepdim=True) + eps) * weight
        if cos is not None:
            yh = y.view(B, T, H, 2, dh // 2).permute(0, 2, 1, 3, 4)
            a_, b_ = yh[..., 0, :], yh[..., 1, :]
            y = torch.stack([a_ * cos - b_ * sin, b_ * cos + a_ * sin], dim=-2).permute(0, 2, 1, 3, 4).reshape(B * T, H * dh)
    hp = H // groups
    out[:, :, slot, :] = y.view(B * T, groups, hp * dh).permute(1, 0, 2).to(out.dtype)
    return out


def fake_attention(q, k, v, out, B, Tq, Tk, H, dh, scale, kv_bias=None):
    qh = q.float().reshape(B, Tq, H, dh).transpose(1, 2)
    kh = k.float().reshape(B, Tk, H, dh).transpose(1, 2)
    vh = v.float().reshape(B, Tk, H, dh).transpose(1, 2)
    p = torch.softmax(qh @ kh.transpose(-1, -2) * scale, -1)
    out.copy_((p @ vh).transpose(1, 2).reshape(B * Tq, H * dh).to(out.dtype))
    return out


class FakeAttn:
    """The attributes UlyssesGroup.self_attention reads off an Attention module."""

    def __init__(self, D, H, dh, g):
        self.heads, self.dim_head, self.inner_dim = H, dh, H * dh
        self.qkv_weight = (torch.randn(3 * H * dh, D, generator=g) / math.sqrt(D)).bfloat16()
        self.qkv_bias = torch.randn(3 * H * dh, generator=g) * 0.1
        self.q_norm = type("N", (), {"weight": 1 + 0.1 * torch.randn(H * dh, generator=g), "eps": 1e-6})()
        self.k_norm = type("N", (), {"weight": 1 + 0.1 * torch.randn(H * dh, generator=g), "eps": 1e-6})()

    def sdpa(self, ws, tag, q, k, v, B, Tq, Tk, kv_bias, heads=None):
        H = self.heads if heads is None else heads
        o = torch.empty(B * Tq, H * self.dim_head, dtype=torch.bfloat16)
        return fake_attention(q, k, v, o, B, Tq, Tk, H, self.dim_head, 1 / math.sqrt(self.dim_head), kv_bias)


def _worker(rank: int, world: int, port: int):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ops.gemm, ops.qknorm_rope_scatter, ops.attention = fake_gemm, fake_scatter, fake_attention
        g = torch.Generator().manual_seed(0)  # identical on both ranks
        # ---------------- token sharding + gather
        uly = parallel.UlyssesGroup(list(range(world)), rank, fused=False)
        T, Tc = 24, 6
        cos, sin = torch.randn(1, 4, T, 64, generator=g), torch.randn(1, 4, T, 64, generator=g)
        m = M.Modality(latent=torch.randn(1, T, 128, generator=g), timesteps=torch.rand(1, T, generator=g),
                       positions=torch.randn(1, 3, T, 2, generator=g), context=torch.randn(1, Tc, 32, generator=g),
                       positional_embeddings=(cos, sin))
        vm, am = uly.shard_inputs(m, None)
        sl = slice(rank * T // world, (rank + 1) * T // world)
        assert am is None and torch.equal(vm.latent, m.latent[:, sl]) and torch.equal(vm.timesteps, m.timesteps[:, sl])
        assert torch.equal(vm.positions, m.positions[:, :, sl]) and vm.context is m.context
        assert torch.equal(vm.positional_embeddings[0], cos[:, :, sl]) and vm.positional_embeddings[0].is_contiguous()
        assert uly.shard_inputs(m, None)[0].positional_embeddings[0] is vm.positional_embeddings[0], "rope slice must be cached"
        one = M.Modality(m.latent, torch.tensor([[0.5]]), m.positions, m.context)
        assert uly.shard_modality(one).timesteps.shape == (1, 1)  # (B, 1) timesteps are not token-sharded
        full, _ = uly.gather_outputs(vm.latent, None)
        assert torch.equal(full, m.latent)
        with pytest.raises(_lib.LtxbError):
            uly.local_slice(25)
        with pytest.raises(_lib.LtxbError):
            uly.shard_inputs(M.Modality(torch.zeros(2, T, 128), m.timesteps, m.positions, m.context), None)

        # ---------------- sequence-parallel self-attention == single-process attention on the same rows
        D, H, dh = 64, 4, 64
        attn = FakeAttn(D, H, dh, g)
        x = torch.randn(T, D, generator=g).bfloat16()
        pe = (torch.cos(cos[:, :, :, :dh // 2]), torch.sin(cos[:, :, :, :dh // 2]))
        ws = Workspace()
        qkv = torch.empty(T, 3 * H * dh, dtype=torch.bfloat16)
        fake_gemm(x, attn.qkv_weight, attn.qkv_bias, qkv)
        one_buf = torch.empty(1, T, 3, H * dh, dtype=torch.bfloat16)  # groups = 1: the un-sharded layout
        fake_scatter(qkv[:, :H * dh], one_buf, 0, 3, 1, 1, T, H, dh, attn.q_norm.weight, 1e-6, pe[0], pe[1])
        fake_scatter(qkv[:, H * dh:2 * H * dh], one_buf, 1, 3, 1, 1, T, H, dh, attn.k_norm.weight, 1e-6, pe[0], pe[1])
        fake_scatter(qkv[:, 2 * H * dh:], one_buf, 2, 3, 1, 1, T, H, dh, None, 0.0)
        want = attn.sdpa(ws, "ref", one_buf[0, :, 0], one_buf[0, :, 1], one_buf[0, :, 2], 1, T, T, None)  # [T, H*dh]
        pe_local = (pe[0][:, :, sl].contiguous(), pe[1][:, :, sl].contiguous())
        back, gcols = uly.self_attention(attn, ws, "t", x[sl], 1, T // world, pe_local)
        assert gcols == (H // world) * dh and back.shape == (world, T // world, gcols)
        got = back.permute(1, 0, 2).reshape(T // world, H * dh)  # what the grouped-A GEMM reads
        assert (got.float() - want[sl].float()).abs().max() < 2e-2
        # and the to_out GEMM over the grouped operand equals the plain one
        w_out = (torch.randn(D, H * dh, generator=g) / 16).bfloat16()
        y1 = fake_gemm(back, w_out, None, torch.empty(T // world, D, dtype=torch.bfloat16), a_group_cols=gcols)
        y2 = fake_gemm(got, w_out, None, torch.empty(T // world, D, dtype=torch.bfloat16))
        assert torch.equal(y1, y2)

        # ---------------- CFG-parallel exchange
        cfgp = parallel.CFGParallel(world, rank)
        assert cfgp.is_cond == (rank == 0) and cfgp.partner == 1 - rank
        mine = torch.full((1, 5, 128), float(rank + 1))
        v_pos, v_neg = cfgp.exchange(mine)
        assert float(v_pos.mean()) == 1.0 and float(v_neg.mean()) == 2.0
    finally:
        dist.destroy_process_group()


def test_two_rank_host_logic():
    mp.spawn(_worker, args=(2, _free_port()), nprocs=2, join=True)


def test_layout_plan():
    assert parallel.plan(1, False) == (False, 1) and parallel.plan(1, True) == (False, 1)
    assert parallel.plan(2, True) == (True, 1) and parallel.plan(2, False) == (False, 2)
    assert parallel.plan(4, True) == (True, 2) and parallel.plan(8, True) == (True, 4) and parallel.plan(8, False) == (False, 8)
    with pytest.raises(ValueError):
        parallel.plan(3, False)
