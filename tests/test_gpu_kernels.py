"""GPU: each C-ABI kernel against a plain PyTorch fp32 reference of the same op (tolerance in each test)."""
import math

import pytest
import torch

import mlx_video_b200  # noqa: F401
from conftest import rel_l2
from mlx_video_b200 import _lib, ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def gemm_case(M, N, K, mode, pair=-1, bn=0, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    a = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    w = (torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device=DEV, generator=g)
    acc = a.float() @ w.float().T + bias
    kw = {}
    if mode == _lib.EPI_BIAS_BF16:
        ref, out = acc, torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    elif mode == _lib.EPI_GELU_BF16:
        ref, out = torch.nn.functional.gelu(acc, approximate="tanh"), torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    elif mode == _lib.EPI_SILU_BF16:
        ref, out = torch.nn.functional.silu(acc), torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    elif mode == _lib.EPI_BIAS_F32:
        ref, out = acc, torch.empty(M, N, device=DEV, dtype=torch.float32)
    else:
        resid = torch.randn(M, N, device=DEV, generator=g)
        div = 4
        gate = torch.randn((M + div - 1) // div, 3 * N, device=DEV, generator=g)
        table = torch.randn(N, device=DEV, generator=g)
        gsl = gate[:, N:2 * N]
        ref = resid + acc * (table + gsl[torch.arange(M, device=DEV) // div])
        out = resid.clone()
        kw = dict(resid=out, gate=gsl, gate_table=table, gate_row_div=div)
    ops.gemm(a, w, bias, out, mode=mode, block_n=bn, cta_pair=pair, **kw)
    torch.cuda.synchronize()
    # fp32 out: summation order only (of the kernel's and of the reference's own fp32 matmul: grows ~sqrt(K)); bf16 out: output rounding
    tol = 2e-5 * max(1.0, (K / 16384) ** 0.5) if out.dtype == torch.float32 else 4e-3
    err = rel_l2(out.float(), ref)
    assert torch.isfinite(out.float()).all() and err <= tol, f"gemm M={M} N={N} K={K} mode={mode} pair={pair} bn={bn}: {err:.3e}"
    return out


# (M, N, K): LTX-2 shapes at T=1280 / Tc=1024 (stream-K territory: ragged last wave), ragged M, tiny M, narrow N
SHAPES = [(1280, 4096, 4096), (1280, 12288, 4096), (1024, 8192, 4096), (1280, 4096, 16384), (1280, 16384, 4096),
          (5184, 4096, 4096), (1000, 1024, 512), (68, 2048, 2048), (2, 6144, 1024), (1280, 128, 4096), (1280, 4096, 128),
          (257, 272, 192), (300, 512, 256), (648, 4096, 4096), (128, 24576, 4096)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_auto_schedule(M, N, K):
    """block_n=0 / cta_pair=-1: the library picks tile, CTA pairing and data-parallel vs stream-K."""
    gemm_case(M, N, K, _lib.EPI_BIAS_F32)


@pytest.mark.parametrize("mode", [_lib.EPI_BIAS_BF16, _lib.EPI_GELU_BF16, _lib.EPI_SILU_BF16, _lib.EPI_RESID_GATE_F32])
@pytest.mark.parametrize("M,N,K", [(1280, 4096, 4096), (520, 784, 512), (1280, 12288, 4096)])
def test_gemm_epilogues(M, N, K, mode):
    gemm_case(M, N, K, mode)
    gemm_case(M, N, K, mode, pair=0)


@pytest.mark.parametrize("pair,bn", [(0, 128), (1, 128), (0, 256), (1, 256), (1, 144), (0, 48), (1, 224)])
def test_gemm_explicit_tiles(pair, bn):
    for M, N, K in [(300, 512, 256), (1280, 4096, 1024), (257, 272, 192)]:
        gemm_case(M, N, K, _lib.EPI_BIAS_F32, pair=pair, bn=bn)


def test_gemm_split_k_matches_data_parallel_and_is_reproducible():
    """An explicit block_n forces the data-parallel schedule; auto picks stream-K for this shape.  Both must
    agree to fp32 summation order, and split-K must be bit-reproducible run to run (fixed k-order fix-up)."""
    for M, N, K in [(1280, 4096, 4096), (1280, 4096, 16384), (1024, 8192, 4096)]:
        sk1 = gemm_case(M, N, K, _lib.EPI_BIAS_F32)
        sk2 = gemm_case(M, N, K, _lib.EPI_BIAS_F32)
        dp = gemm_case(M, N, K, _lib.EPI_BIAS_F32, pair=1, bn=256)
        assert torch.equal(sk1, sk2), "stream-K result differs between two runs"
        assert rel_l2(sk1, dp) < 2e-5
    # many back-to-back launches reuse the same scratch + counters
    outs = [gemm_case(1280, 4096, 4096, _lib.EPI_GELU_BF16, seed=3) for _ in range(5)]
    assert all(torch.equal(outs[0], o) for o in outs[1:])


# Sequence-parallel shards: M = T / ranks rows per GPU (1280 / 8 = 160, 1280 / 4 = 320, 5184 / 8 = 648, 68 audio tokens)
# against the full LTX-2 weight matrices — the weight-streaming regime where the contiguous stream-K schedule applies.
SMALL_M_SHAPES = [(160, 4096, 4096), (160, 12288, 4096), (160, 16384, 4096), (160, 4096, 16384), (320, 4096, 4096),
                  (320, 16384, 4096), (640, 12288, 4096), (68, 2048, 2048), (128, 4096, 4096), (1, 4096, 1024), (200, 272, 1024)]


@pytest.mark.parametrize("M,N,K", SMALL_M_SHAPES)
def test_gemm_contiguous_stream_k(M, N, K):
    """cta_pair 2 / 3 force the contiguous stream-K schedule (single-CTA / pair tiles); -1 lets the library choose it.
    All must agree with the fp32 reference and with each other to summation order, and be bit-reproducible."""
    auto = gemm_case(M, N, K, _lib.EPI_BIAS_F32)
    for pair in (2, 3):
        a = gemm_case(M, N, K, _lib.EPI_BIAS_F32, pair=pair)
        b = gemm_case(M, N, K, _lib.EPI_BIAS_F32, pair=pair)
        assert torch.equal(a, b), f"contiguous stream-K (cta_pair={pair}) differs between two runs"
        assert rel_l2(a, auto) < 2e-5
    assert torch.equal(auto, gemm_case(M, N, K, _lib.EPI_BIAS_F32))


@pytest.mark.parametrize("mode", [_lib.EPI_BIAS_BF16, _lib.EPI_GELU_BF16, _lib.EPI_RESID_GATE_F32])
def test_gemm_contiguous_stream_k_epilogues(mode):
    for M, N, K in [(160, 4096, 4096), (160, 16384, 4096), (160, 4096, 16384), (648, 4096, 4096)]:
        for pair in (-1, 2, 3):
            gemm_case(M, N, K, mode, pair=pair)
    # back-to-back launches reuse the parked-partial slots and the arrival counters
    outs = [gemm_case(160, 4096, 16384, _lib.EPI_GELU_BF16, pair=3, seed=5) for _ in range(6)]
    assert all(torch.equal(outs[0], o) for o in outs[1:])


# The weight-streaming kernel for few rows (gemm_small_m.cu; cta_pair=4 forces it, block_n = number of k-range splits):
# sequence-parallel shards, audio tokens, AdaLN rows, ragged N (not a multiple of the 256-row weight tile), two-MMA
# token ranges (M > 256), K down to one k-block per split.
WS_SHAPES = [(160, 4096, 4096), (160, 12288, 4096), (160, 16384, 4096), (160, 4096, 16384), (68, 2048, 2048), (136, 2048, 8192),
             (1, 4096, 1024), (2, 36864, 4096), (16, 256, 64), (200, 272, 1024), (256, 4096, 4096), (255, 1040, 512),
             (320, 4096, 4096), (320, 16384, 4096), (512, 4096, 4096), (257, 784, 256), (500, 12288, 4096),
             (200, 1024, 27648), (136, 512, 13824), (200, 128, 3456)]  # + VAE decoder convolutions in row chunks: K = 27 * C


@pytest.mark.parametrize("M,N,K", WS_SHAPES)
def test_gemm_small_m_kernel(M, N, K):
    """Forced (cta_pair=4) with the library's own split choice, with no split, and with 2 / 3 / 7 k-range splits: all agree
    with the fp32 reference, with each other to summation order, and each is bit-reproducible run to run."""
    auto = gemm_case(M, N, K, _lib.EPI_BIAS_F32, pair=4)
    assert torch.equal(auto, gemm_case(M, N, K, _lib.EPI_BIAS_F32, pair=4))
    for splits in (1, 2, 3, 7):
        a = gemm_case(M, N, K, _lib.EPI_BIAS_F32, pair=4, bn=splits)
        b = gemm_case(M, N, K, _lib.EPI_BIAS_F32, pair=4, bn=splits)
        assert torch.equal(a, b), f"small-M kernel with {splits} splits differs between two runs"
        assert rel_l2(a, auto) < 2e-5 * max(1.0, (K / 16384) ** 0.5)  # summation order only
    if M <= 512:  # what the library picks on its own for few rows
        assert torch.equal(auto, gemm_case(M, N, K, _lib.EPI_BIAS_F32))


@pytest.mark.parametrize("mode", [_lib.EPI_BIAS_BF16, _lib.EPI_GELU_BF16, _lib.EPI_SILU_BF16, _lib.EPI_RESID_GATE_F32])
def test_gemm_small_m_kernel_epilogues(mode):
    for M, N, K in [(160, 4096, 4096), (160, 16384, 4096), (160, 4096, 16384), (68, 2048, 2048), (320, 4096, 4096), (3, 784, 512)]:
        for splits in (0, 1, 4):
            gemm_case(M, N, K, mode, pair=4, bn=splits)
    # back-to-back launches reuse the parked chunks and the arrival / departure counters
    outs = [gemm_case(160, 4096, 16384, mode, pair=4, seed=5) for _ in range(6)]
    assert all(torch.equal(outs[0], o) for o in outs[1:])


def test_gemm_small_m_kernel_group_major_rows_and_strides():
    """The A operand as the Ulysses gather all-to-all delivers it ([K / g][M][g], read in place through a 3-D map), a
    gate looked up through a row index, and outputs / residuals that are column slices of wider tensors."""
    g = torch.Generator(device=DEV).manual_seed(11)
    for M, N, K, gc in [(320, 4096, 4096, 1024), (160, 4096, 4096, 512)]:  # 4 / 8 sequence-parallel ranks (two MMAs per k-slice at 320)
        a = torch.randn(M, K, device=DEV, generator=g).bfloat16()
        w = (torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K)).bfloat16()
        bias = torch.randn(N, device=DEV, generator=g)
        a_grouped = a.view(M, K // gc, gc).permute(1, 0, 2).contiguous()
        ref = a.float() @ w.float().T + bias
        for pair in (4, -1):
            out = torch.empty(M, N, device=DEV, dtype=torch.float32)
            ops.gemm(a_grouped, w, bias, out, mode=_lib.EPI_BIAS_F32, a_group_cols=gc, cta_pair=pair)
            assert rel_l2(out, ref) < 2e-5
    wide = torch.randn(M, 3 * N, device=DEV, generator=g)
    resid = wide[:, N:2 * N]
    before = wide.clone()
    want = resid.clone()
    idx = torch.randint(0, 5, (M,), device=DEV, generator=g, dtype=torch.int32)
    gate = torch.randn(5, N, device=DEV, generator=g)
    want = want + ref * gate[idx.long()]
    ops.gemm(a, w, bias, resid, mode=_lib.EPI_RESID_GATE_F32, resid=resid, gate=gate, gate_row_index=idx, cta_pair=4)
    assert rel_l2(resid, want) < 2e-5
    assert torch.equal(wide[:, :N], before[:, :N]) and torch.equal(wide[:, 2 * N:], before[:, 2 * N:])


def test_gemm_small_m_kernel_const_weights_flag():
    """LTXB_GEMM_CONST_W (weights requested before the programmatic-launch wait) gives the same bits as without it, also when
    a split has fewer k-blocks than pipeline stages."""
    g = torch.Generator(device=DEV).manual_seed(17)
    for M, N, K, splits in [(160, 4096, 4096, 0), (160, 4096, 256, 2), (68, 2048, 128, 1), (320, 12288, 4096, 0), (16, 512, 64, 1)]:
        a = torch.randn(M, K, device=DEV, generator=g).bfloat16()
        w = (torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K)).bfloat16()
        bias = torch.randn(N, device=DEV, generator=g)
        outs = []
        for const_w in (False, True, True):
            out = torch.empty(M, N, device=DEV, dtype=torch.float32)
            ops.gemm(a, w, bias, out, mode=_lib.EPI_BIAS_F32, cta_pair=4, block_n=splits, const_w=const_w)
            outs.append(out)
        torch.cuda.synchronize()
        assert rel_l2(outs[0], a.float() @ w.float().T + bias) < 2e-5
        assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])


def test_gemm_small_m_kernel_graph_replay():
    """Captured and replayed back to back (the counters are re-armed inside the kernel), interleaved with the big-tile
    split-K kernel that shares the workspace."""
    g = torch.Generator(device=DEV).manual_seed(13)
    M, K = 160, 4096
    a = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    w1 = (torch.randn(16384, K, device=DEV, generator=g) / math.sqrt(K)).bfloat16()
    w2 = (torch.randn(4096, 16384, device=DEV, generator=g) / 128).bfloat16()
    big_a = torch.randn(1280, K, device=DEV, generator=g).bfloat16()
    h = torch.empty(M, 16384, device=DEV, dtype=torch.bfloat16)
    y = torch.empty(M, 4096, device=DEV, dtype=torch.float32)
    big = torch.empty(1280, 4096, device=DEV, dtype=torch.float32)

    def run():
        # const_w: the weights start streaming before the producer of `h` has finished; `h` itself must still wait for it
        ops.gemm(a, w1, None, h, mode=_lib.EPI_GELU_BF16, const_w=True)
        ops.gemm(h, w2, None, y, mode=_lib.EPI_BIAS_F32, const_w=True)
        ops.gemm(big_a, w2[:, :K], None, big, mode=_lib.EPI_BIAS_F32)

    run()
    torch.cuda.synchronize()
    want_y, want_big = y.clone(), big.clone()
    ref = torch.nn.functional.gelu(a.float() @ w1.float().T, approximate="tanh").bfloat16().float() @ w2.float().T
    assert rel_l2(want_y, ref) < 3e-3
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        run()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            run()
            run()
        for _ in range(4):
            y.zero_()
            big.zero_()
            graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(y, want_y) and torch.equal(big, want_big)


def test_gemm_rejects_bad_shapes():
    a = torch.zeros(64, 96, device=DEV, dtype=torch.bfloat16)
    w = torch.zeros(64, 96, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(_lib.LtxbError, match="multiple of 64"):
        ops.gemm(a, w, None, torch.empty(64, 64, device=DEV, dtype=torch.bfloat16))


@pytest.mark.parametrize("B,Tq,Tk,H,dh,bias", [(1, 128, 128, 1, 128, False), (1, 256, 384, 2, 128, False), (2, 200, 72, 3, 128, False),
                                                (1, 1280, 1280, 32, 128, False), (2, 320, 1024, 4, 128, True), (1, 128, 128, 2, 64, False),
                                                (1, 300, 68, 32, 64, False), (1, 68, 1280, 32, 64, False), (1, 1280, 1024, 8, 128, False)])
def test_attention(B, Tq, Tk, H, dh, bias):
    g = torch.Generator(device=DEV).manual_seed(7)
    D = H * dh
    q = torch.randn(B * Tq, D, device=DEV, generator=g).bfloat16()
    kv = torch.randn(B * Tk, 2 * D, device=DEV, generator=g).bfloat16()
    k, v = kv[:, :D], kv[:, D:]
    kb = None
    if bias:
        kb = torch.zeros(B, Tk, device=DEV)
        kb[:, Tk - Tk // 3:] = -1e9
    out = torch.empty(B * Tq, D, device=DEV, dtype=torch.bfloat16)
    ops.attention(q, k, v, out, B, Tq, Tk, H, dh, 1.0 / math.sqrt(dh), kb)
    qh = q.float().view(B, Tq, H, dh).transpose(1, 2)
    kh = k.float().reshape(B, Tk, H, dh).transpose(1, 2)
    vh = v.float().reshape(B, Tk, H, dh).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) / math.sqrt(dh)
    if kb is not None:
        s = s + kb[:, None, None, :]
    ref = (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(B * Tq, D)
    err = rel_l2(out.float(), ref)
    assert err <= 8e-3, f"attention rel_l2 {err:.3e}"  # bf16 P and bf16 output rounding
    # KAT (iv): softmax rows sum to 1 -> attention of constant V returns that constant
    ones = torch.ones_like(v)
    ops.attention(q, k, ones, out, B, Tq, Tk, H, dh, 1.0 / math.sqrt(dh), kb)
    assert float((out.float() - 1).abs().max()) <= 8e-3


def _attention_ref_fp32(q, k, v, B, Tq, Tk, H, dh):
    """Plain PyTorch fp32 reference of attention.py:13-53, one (batch, head) at a time so the (Tq, Tk) score matrix of
    the 14 080-token config (793 MB in fp32) is the largest temporary."""
    D = H * dh
    out = torch.empty(B * Tq, D, device=q.device, dtype=torch.float32)
    q4, k4, v4 = q.view(B, Tq, H, dh), k.reshape(B, Tk, H, dh), v.reshape(B, Tk, H, dh)
    for b in range(B):
        for h in range(H):
            s = (q4[b, :, h].float() @ k4[b, :, h].float().T) / math.sqrt(dh)
            out.view(B, Tq, H, dh)[b, :, h] = torch.softmax(s, -1) @ v4[b, :, h].float()
            del s
    return out


# BASELINE.json configs[2..4]: dev pipeline B=2 x 5184 tokens (CFG batch) self- and text-cross-attention, the 14 080-token
# long config, and the audio<->video cross-attentions of the joint model (5184 video x 68 audio tokens, 64-wide heads).
# These are the shapes bench.py --workload dev|av|long times: job decode past one wave, key split of the ragged wave,
# single-64-key last tiles (68 = 64 + 4, 5184 = 40 * 128 + 64).
@pytest.mark.parametrize("B,Tq,Tk,H,dh", [(2, 5184, 5184, 32, 128), (1, 14080, 14080, 32, 128), (2, 5184, 1024, 32, 128),
                                           (1, 5184, 68, 32, 64), (1, 68, 5184, 32, 64), (1, 1280, 1024, 32, 128),
                                           (1, 3520, 3520, 32, 128), (1, 648, 5184, 4, 128)])
def test_attention_baseline_config_shapes(B, Tq, Tk, H, dh):
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device=DEV).manual_seed(11)
    D = H * dh
    q = torch.randn(B * Tq, D, device=DEV, generator=g).bfloat16()
    kv = torch.randn(B * Tk, 2 * D, device=DEV, generator=g).bfloat16()
    k, v = kv[:, :D], kv[:, D:]
    out = torch.empty(B * Tq, D, device=DEV, dtype=torch.bfloat16)
    ops.attention(q, k, v, out, B, Tq, Tk, H, dh, 1.0 / math.sqrt(dh))
    ref = _attention_ref_fp32(q, k, v, B, Tq, Tk, H, dh)
    err = rel_l2(out.float(), ref)
    assert torch.isfinite(out.float()).all() and err <= 8e-3, f"attention B={B} Tq={Tq} Tk={Tk} dh={dh}: rel_l2 {err:.3e}"
    # per-row check as well: no single query row may be off (a wrong key range in ONE job would hide in the global norm)
    row_err = (out.float() - ref).norm(dim=1) / ref.norm(dim=1).clamp_min(1e-20)
    assert float(row_err.max()) <= 5e-2, f"worst row rel_l2 {float(row_err.max()):.3e} at row {int(row_err.argmax())}"
    out2 = torch.empty_like(out)
    ops.attention(q, k, v, out2, B, Tq, Tk, H, dh, 1.0 / math.sqrt(dh))
    assert torch.equal(out, out2), "attention is not deterministic"


@pytest.mark.parametrize("B,Tq,Tk,H,dh,parts", [(1, 68, 5184, 32, 64, 8), (1, 68, 1296, 32, 64, 4), (2, 37, 300, 4, 128, 3), (1, 130, 648, 8, 64, 2)])
def test_attention_partial_and_merge(B, Tq, Tk, H, dh, parts):
    """Video->audio attention under sequence parallelism: un-normalised attention over key SLICES (ltxb_attention_partial)
    merged by log-sum-exp (ltxb_attention_merge) == attention over all keys (attention.py:13-53)."""
    g = torch.Generator(device=DEV).manual_seed(13)
    D = H * dh
    q = torch.randn(B * Tq, D, device=DEV, generator=g).bfloat16()
    k = (torch.randn(B, Tk, D, device=DEV, generator=g) * 1.5).bfloat16()
    v = torch.randn(B, Tk, D, device=DEV, generator=g).bfloat16()
    n = ops.attention_partial_floats(B, Tq, H, dh)
    blocks = torch.zeros(parts, n, device=DEV)
    Tl = Tk // parts
    for i in range(parts):
        ks, vs = k[:, i * Tl:(i + 1) * Tl].reshape(B * Tl, D).contiguous(), v[:, i * Tl:(i + 1) * Tl].reshape(B * Tl, D).contiguous()
        ops.attention_partial(q, ks, vs, blocks[i], B, Tq, Tl, H, dh, 1.0 / math.sqrt(dh))
    out = torch.empty(B * Tq, D, device=DEV, dtype=torch.bfloat16)
    ops.attention_merge(blocks, out, B, Tq, H, dh)
    used = Tl * parts
    ref = _attention_ref_fp32(q, k[:, :used].reshape(B * used, D), v[:, :used].reshape(B * used, D), B, Tq, used, H, dh)
    err = rel_l2(out.float(), ref)
    assert torch.isfinite(out.float()).all() and err <= 8e-3, f"partial+merge rel_l2 {err:.3e}"
    # one part == plain attention, to the bf16 rounding of the output
    whole = torch.empty_like(out)
    ops.attention(q, k[:, :used].reshape(B * used, D), v[:, :used].reshape(B * used, D), whole, B, Tq, used, H, dh, 1.0 / math.sqrt(dh))
    assert rel_l2(out.float(), whole.float()) <= 8e-3


def test_rmsnorm_layernorm_modulate():
    g = torch.Generator(device=DEV).manual_seed(1)
    for R, D in [(1280, 4096), (68, 2048), (37, 512), (5, 1032)]:
        x = torch.randn(R, D, device=DEV, generator=g) * 3
        out = torch.empty(R, D, device=DEV, dtype=torch.bfloat16)
        ops.rmsnorm_modulate(x, out, 1e-6)
        ref = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6)
        assert rel_l2(out.float(), ref) <= 4e-3
        mod = torch.randn(3, 6 * D, device=DEV, generator=g) * 0.1
        table = torch.randn(6, D, device=DEV, generator=g) * 0.1
        idx = torch.randint(0, 3, (R,), device=DEV, generator=g, dtype=torch.int32)
        ops.rmsnorm_modulate(x, out, 1e-6, mod=mod, scale_off=4 * D, shift_off=3 * D, table_scale=table[4], table_shift=table[3], row_index=idx)
        sc = table[4] + mod[idx.long(), 4 * D:5 * D]
        sh = table[3] + mod[idx.long(), 3 * D:4 * D]
        assert rel_l2(out.float(), ref * (1 + sc) + sh) <= 4e-3
        emb = torch.randn(3, D, device=DEV, generator=g) * 0.1
        ops.layernorm_modulate(x, out, 1e-6, emb, table[1], table[0], row_index=idx)
        ln = torch.nn.functional.layer_norm(x, (D,), eps=1e-6)
        assert rel_l2(out.float(), ln * (1 + table[1] + emb[idx.long()]) + table[0] + emb[idx.long()]) <= 4e-3


def test_residual_rmsnorm_modulate():
    """x += y * gate and the norm of the updated row in one pass == the two separate steps (fp32 torch reference):
    x to fp32 rounding (1e-6), the normalised bf16 output to its own rounding (4e-3)."""
    g = torch.Generator(device=DEV).manual_seed(5)
    for R, D in [(1280, 4096), (68, 2048), (37, 512), (5, 1032), (3, 8192)]:
        x0 = torch.randn(R, D, device=DEV, generator=g) * 3
        y = torch.randn(R, D, device=DEV, generator=g).bfloat16()
        mod = torch.randn(3, 6 * D, device=DEV, generator=g) * 0.1
        table = torch.randn(6, D, device=DEV, generator=g) * 0.1
        idx = torch.randint(0, 3, (R,), device=DEV, generator=g, dtype=torch.int32)
        out = torch.empty(R, D, device=DEV, dtype=torch.bfloat16)
        norm = lambda t: t * torch.rsqrt(t.pow(2).mean(-1, keepdim=True) + 1e-6)  # noqa: E731
        # gated residual, plain norm (attn1 out-projection -> attn2's norm)
        x = x0.clone()
        ops.residual_rmsnorm_modulate(x, y, out, 1e-6, mod=mod, gate_off=2 * D, table_gate=table[2], row_index=idx)
        xr = x0 + y.float() * (table[2] + mod[idx.long(), 2 * D:3 * D])
        assert rel_l2(x, xr) <= 1e-6 and rel_l2(out.float(), norm(xr)) <= 4e-3
        # ungated residual, modulated norm (attn2 out-projection -> the FFN's norm); row divisor instead of an index
        x = x0.clone()
        div = 2
        mod_b = torch.randn((R + div - 1) // div, 6 * D, device=DEV, generator=g) * 0.1
        ops.residual_rmsnorm_modulate(x, y, out, 1e-6, mod=mod_b, scale_off=4 * D, shift_off=3 * D, table_scale=table[4],
                                      table_shift=table[3], row_div=div)
        xr = x0 + y.float()
        rows = torch.arange(R, device=DEV) // div
        ref = norm(xr) * (1 + table[4] + mod_b[rows, 4 * D:5 * D]) + table[3] + mod_b[rows, 3 * D:4 * D]
        assert rel_l2(x, xr) <= 1e-6 and rel_l2(out.float(), ref) <= 4e-3
        # nothing but the add
        x = x0.clone()
        ops.residual_rmsnorm_modulate(x, y, out, 1e-6)
        assert rel_l2(x, xr) <= 1e-6 and rel_l2(out.float(), norm(xr)) <= 4e-3


def test_qknorm_rope_and_small_kernels():
    g = torch.Generator(device=DEV).manual_seed(2)
    for B, T, H, dh in [(1, 320, 32, 128), (2, 68, 32, 64), (1, 77, 4, 128)]:
        D = H * dh
        qkv = torch.randn(B * T, 3 * D, device=DEV, generator=g).bfloat16()
        x = qkv[:, D:2 * D]
        x0 = x.float().clone()
        wgt = 1 + 0.1 * torch.randn(D, device=DEV, generator=g)
        ang = torch.rand(B, H, T, dh // 2, device=DEV, generator=g) * 6.28
        cos, sin = torch.cos(ang), torch.sin(ang)
        ops.qknorm_rope(x, B, T, H, dh, wgt, 1e-6, cos, sin)
        n = x0 * torch.rsqrt(x0.pow(2).mean(-1, keepdim=True) + 1e-6) * wgt
        nh = n.view(B, T, H, 2, dh // 2).permute(0, 2, 1, 3, 4)
        a_, b_ = nh[..., 0, :], nh[..., 1, :]
        rot = torch.stack([a_ * cos - b_ * sin, b_ * cos + a_ * sin], dim=-2).permute(0, 2, 1, 3, 4).reshape(B * T, D)
        assert rel_l2(x.float(), rot) <= 4e-3
        assert torch.equal(qkv[:, :D].float(), qkv[:, :D].float()) and torch.isfinite(qkv.float()).all()
    # segments: several column slices of one buffer, each with its own weight, in one launch == one call per slice
    for B, T, H, dh, n_seg, stride, rope in [(1, 320, 32, 128, 2, 4096, True), (2, 40, 4, 64, 5, 512, False)]:
        D = H * dh
        buf = torch.randn(B * T, n_seg * stride + 64, device=DEV, generator=g).bfloat16()
        ref = buf.clone()
        wgt = 1 + 0.1 * torch.randn(n_seg, D, device=DEV, generator=g)
        w2 = None if rope else 1 + 0.1 * torch.randn(n_seg, D, device=DEV, generator=g)
        ang = torch.rand(B, H, T, dh // 2, device=DEV, generator=g) * 6.28
        cs = (torch.cos(ang), torch.sin(ang)) if rope else (None, None)
        ops.qknorm_rope_segments(buf, n_seg, stride, B, T, H, dh, wgt, 1e-6, cs[0], cs[1], weight2=w2)
        for s_ in range(n_seg):
            sl = ref[:, s_ * stride:s_ * stride + D]
            ops.qknorm_rope(sl, B, T, H, dh, wgt[s_] if w2 is None else (wgt[s_] * w2[s_]).contiguous(), 1e-6, cs[0], cs[1])
        assert rel_l2(buf.float(), ref.float()) <= (0.0 if w2 is None else 4e-3)
    t = torch.tensor([0.0, 0.05, 0.421875, 0.725, 1.0], device=DEV)
    feat = torch.empty(5, 256, device=DEV, dtype=torch.bfloat16)
    ops.timestep_embed(t, 1000.0, 256, feat)
    f = torch.exp(-math.log(10000.0) * torch.arange(128, device=DEV) / 128)
    arg = (t * 1000.0)[:, None] * f[None]
    assert rel_l2(feat.float(), torch.cat([torch.cos(arg), torch.sin(arg)], -1)) <= 4e-3
    # timestep grouping: every token's value is recoverable through (values, index)
    ts = torch.tensor([0.725, 0.0, 0.725, 0.5, 0.0, -0.0, 0.5] * 300, device=DEV)
    values, index, count = ops.timestep_groups(ts, 16)
    assert int(count) == 3 and torch.equal(values[index.long()], ts.abs() * torch.sign(ts).abs())
    # sampler step
    x = torch.randn(96, 128, device=DEV, generator=g)
    vp, vn = torch.randn(96, 128, device=DEV, generator=g), torch.randn(96, 128, device=DEV, generator=g)
    mask = (torch.rand(96, device=DEV, generator=g) > 0.3).float()
    clean = torch.randn(96, 128, device=DEV, generator=g)
    x_ref = x.clone()
    v = vp + 3.5 * (vp - vn)
    x0 = (x_ref - 0.725 * v) * mask[:, None] + clean * (1 - mask[:, None])
    want = x0 + 0.421875 * (x_ref - x0) / 0.725
    ops.euler_step(x, vp, 0.725, 0.421875, v_neg=vn, cfg_scale=4.5, mask=mask, clean=clean)
    assert rel_l2(x, want) <= 1e-6
