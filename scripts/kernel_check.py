"""GPU bring-up checks for the C-ABI kernels against plain torch references (run under gpurun).

Usage: python scripts/kernel_check.py <group> [args]; each group prints JSON lines and exits non-zero on
a failed check.  Groups are run in separate processes (a trapped kernel poisons its CUDA context).
"""
import json
import math
import os
import sys
import time

import torch

sys.path.insert(0, ".")
import mlx_video_b200  # noqa: E402,F401
from mlx_video_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
FAIL = 0


def emit(**kw):
    print(json.dumps(kw), flush=True)


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def check(name, got, ref, tol, **extra):
    global FAIL
    err = rel_l2(got.float(), ref.float())
    mx = float((got.float() - ref.float()).abs().max())
    ok = bool(err <= tol) and bool(torch.isfinite(got.float()).all())
    if not ok:
        FAIL += 1
    emit(check=name, ok=ok, rel_l2=err, max_abs=mx, tol=tol, **extra)
    return ok


def time_fn(fn, iters=20, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def err_map(got, ref, bm=128, bn=64):
    """Coarse map of which (row-block, col-block) tiles are wrong — tells descriptor bugs apart."""
    d = (got.float() - ref.float()).abs()
    M, N = d.shape
    rows = []
    for i in range(0, min(M, 4 * bm), bm):
        rows.append([round(float(d[i:i + bm, j:j + bn].max()), 3) for j in range(0, min(N, 8 * bn), bn)])
    return rows


def gemm_case(M, N, K, mode, pair, bn, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    a = torch.randn(M, K, device=dev, generator=g).bfloat16()
    w = (torch.randn(N, K, device=dev, generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device=dev, generator=g)
    acc = a.float() @ w.float().T + bias
    kw = {}
    if mode == _lib.EPI_BIAS_BF16:
        ref, out = acc, torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    elif mode == _lib.EPI_GELU_BF16:
        ref, out = torch.nn.functional.gelu(acc, approximate="tanh"), torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    elif mode == _lib.EPI_SILU_BF16:
        ref, out = torch.nn.functional.silu(acc), torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    elif mode == _lib.EPI_BIAS_F32:
        ref, out = acc, torch.empty(M, N, device=dev, dtype=torch.float32)
    else:
        resid = torch.randn(M, N, device=dev, generator=g)
        div = 4
        gate = torch.randn((M + div - 1) // div, 3 * N, device=dev, generator=g)
        table = torch.randn(N, device=dev, generator=g)
        gsl = gate[:, N:2 * N]
        grow = torch.arange(M, device=dev) // div
        ref = resid + acc * (table + gsl[grow])
        out = resid.clone()
        kw = dict(resid=out, gate=gsl, gate_table=table, gate_row_div=div)
    ops.gemm(a, w, bias, out, mode=mode, block_n=bn, cta_pair=pair, **kw)
    torch.cuda.synchronize()
    tol = 1e-5 if out.dtype == torch.float32 else 4e-3
    ok = check("gemm", out, ref, tol, M=M, N=N, K=K, mode=mode, pair=pair, bn=bn)
    if not ok:
        emit(err_map=err_map(out, ref), note="rows=128-row blocks, cols=64-col blocks")
    return ok


def group_gemm_correct():
    first = gemm_case(256, 256, 128, _lib.EPI_BIAS_F32, 0, 128)
    gemm_case(256, 256, 128, _lib.EPI_BIAS_F32, 1, 128)
    if not first:
        return
    for pair in (0, 1):
        for (M, N, K, bn) in [(128, 128, 64, 128), (300, 512, 256, 256), (1280, 4096, 4096, 0), (1280, 4096, 4096, 144),
                              (1000, 1024, 512, 160), (68, 2048, 2048, 0), (5184, 4096, 1024, 0), (2, 6144, 1024, 0),
                              (1280, 128, 4096, 0), (1280, 4096, 128, 0), (257, 272, 192, 48)]:
            gemm_case(M, N, K, _lib.EPI_BIAS_F32, pair, bn)
        for mode in (_lib.EPI_BIAS_BF16, _lib.EPI_GELU_BF16, _lib.EPI_SILU_BF16, _lib.EPI_RESID_GATE_F32):
            gemm_case(520, 768, 512, mode, pair, 0)
            gemm_case(520, 784, 512, mode, pair, 112)


def group_gemm_perf():
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    shapes = [(1280, 4096, 4096), (1280, 12288, 4096), (1280, 16384, 4096), (1280, 4096, 16384),
              (5184, 4096, 4096), (5184, 16384, 4096), (5184, 4096, 16384), (8192, 8192, 8192)]
    for (M, N, K) in shapes:
        a = torch.randn(M, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
        bias = torch.randn(N, device=dev)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        fl = 2.0 * M * N * K
        med, best = time_fn(lambda: torch.matmul(a, w.T, out=out), flush=flush)
        emit(perf="cublas", M=M, N=N, K=K, ms=med, tflops=fl / med / 1e9, best_tflops=fl / best / 1e9)
        for pair in (0, 1):
            for bn in (0, 128, 144, 160, 192, 208, 224, 256):
                try:
                    med, best = time_fn(lambda: ops.gemm(a, w, bias, out, block_n=bn, cta_pair=pair), flush=flush, iters=10)
                    emit(perf="ltxb", M=M, N=N, K=K, pair=pair, bn=bn, ms=med, tflops=fl / med / 1e9,
                         best_tflops=fl / best / 1e9)
                except Exception as e:  # noqa: BLE001
                    emit(perf="ltxb", M=M, N=N, K=K, pair=pair, bn=bn, error=str(e)[:200])
                    raise


def attn_ref(q, k, v, B, Tq, Tk, H, dh, scale, bias=None):
    qf = q.float().reshape(B, Tq, H, dh).transpose(1, 2)
    kf = k.float().reshape(B, Tk, H, dh).transpose(1, 2)
    vf = v.float().reshape(B, Tk, H, dh).transpose(1, 2)
    s = qf @ kf.transpose(-1, -2) * scale
    if bias is not None:
        s = s + bias[:, None, None, :]
    p = torch.softmax(s, dim=-1)
    return (p @ vf).transpose(1, 2).reshape(B * Tq, H * dh)


def attn_case(B, Tq, Tk, H, dh, kind="rand", use_bias=False, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    D = H * dh
    qkv = torch.randn(B * Tq, 3 * D, device=dev, generator=g).bfloat16()  # strided views like the fused QKV buffer
    q = qkv[:, :D]
    if Tk == Tq:
        k, v = qkv[:, D:2 * D], qkv[:, 2 * D:]
    else:
        kv = torch.randn(B * Tk, 2 * D, device=dev, generator=g).bfloat16()
        k, v = kv[:, :D], kv[:, D:]
    if kind == "uniform":  # K = 0 -> uniform softmax -> O = mean(V): isolates the P.V contraction
        k = torch.zeros_like(k)
    if kind == "onehot":  # huge logits on one key -> O = V[that key]: isolates Q.K^T + indexing
        q = q * 8
    if kind == "drift":  # logits grow with the key index: the running maximum moves in every tile (O rescale path)
        u = torch.randn(1, D, device=dev, generator=g)
        ramp = (torch.arange(Tk, device=dev).float() / Tk).repeat(B)[:, None]
        q = (u + 0.1 * q.float()).bfloat16()
        k = (u * ramp * 6.0 + 0.1 * k.float()).bfloat16()
    bias = None
    if use_bias:
        bias = torch.where(torch.rand(B, Tk, device=dev, generator=g) < 0.3, -1e9, 0.0).float().contiguous()
        bias[:, 0] = 0.0
    out = torch.empty(B * Tq, D, device=dev, dtype=torch.bfloat16)
    scale = 1.0 / math.sqrt(dh)
    ops.attention(q, k, v, out, B, Tq, Tk, H, dh, scale, bias)
    torch.cuda.synchronize()
    ref = attn_ref(q, k, v, B, Tq, Tk, H, dh, scale, bias)
    ok = check("attention", out, ref, 8e-3, B=B, Tq=Tq, Tk=Tk, H=H, dh=dh, kind=kind, bias=use_bias)
    if not ok:
        emit(err_map=err_map(out, ref, bm=128, bn=64))
    return ok


def group_attn_correct():
    attn_case(1, 128, 128, 1, 128, "uniform")
    attn_case(1, 128, 128, 1, 128, "rand")
    attn_case(1, 128, 128, 1, 128, "onehot")
    attn_case(1, 256, 384, 2, 128, "rand")
    attn_case(2, 200, 72, 3, 128, "rand")
    attn_case(1, 256, 128, 1, 128, "uniform")
    attn_case(1, 256, 256, 1, 128, "onehot")
    attn_case(1, 384, 512, 2, 128, "rand")
    attn_case(1, 1280, 1280, 32, 128, "rand")
    attn_case(1, 1280, 1280, 32, 128, "drift")
    attn_case(1, 640, 1024, 2, 128, "drift")
    attn_case(1, 1280, 1024, 32, 128, "rand")
    attn_case(2, 5184, 5184, 4, 128, "rand")
    attn_case(2, 320, 1024, 4, 128, "rand", use_bias=True)
    attn_case(1, 5184, 68, 32, 64, "rand")
    attn_case(1, 1000, 1000, 5, 64, "drift")
    attn_case(1, 128, 128, 1, 64, "uniform")
    attn_case(1, 128, 128, 2, 64, "rand")
    attn_case(1, 300, 68, 32, 64, "rand")
    attn_case(1, 68, 1280, 32, 64, "rand")


def group_attn_perf():
    for (B, T, Tk, H, dh) in [(1, 1280, 1280, 32, 128), (1, 1280, 1024, 32, 128), (2, 5184, 5184, 32, 128),
                              (2, 5184, 1024, 32, 128), (1, 14080, 14080, 32, 128), (1, 5184, 68, 32, 64),
                              (1, 1280, 1280, 4, 128)]:
        D = H * dh
        q = torch.randn(B * T, D, device=dev).bfloat16()
        k = torch.randn(B * Tk, D, device=dev).bfloat16()
        v = torch.randn(B * Tk, D, device=dev).bfloat16()
        out = torch.empty(B * T, D, device=dev, dtype=torch.bfloat16)
        fl = 4.0 * B * H * T * Tk * dh
        med, best = time_fn(lambda: ops.attention(q, k, v, out, B, T, Tk, H, dh, 1 / math.sqrt(dh)), iters=10)
        emit(perf="ltxb_attn", B=B, T=T, Tk=Tk, H=H, dh=dh, us=round(med * 1e3, 1), tflops=round(fl / med / 1e9))
        if os.environ.get("LTXB_SKIP_SDPA"):
            continue
        q4, k4, v4 = (t.reshape(B, -1, H, dh).transpose(1, 2) for t in (q, k, v))
        med, best = time_fn(lambda: torch.nn.functional.scaled_dot_product_attention(q4, k4, v4), iters=10)
        emit(perf="torch_sdpa", B=B, T=T, Tk=Tk, H=H, dh=dh, us=round(med * 1e3, 1), tflops=round(fl / med / 1e9))


def group_elementwise():
    g = torch.Generator(device=dev).manual_seed(1)
    B, T, D = 2, 300, 4096
    R = B * T
    x = torch.randn(R, D, device=dev, generator=g) * 3
    mod = torch.randn(B, 6 * D, device=dev, generator=g) * 0.5
    table = torch.randn(6, D, device=dev, generator=g) * 0.1
    out = torch.empty(R, D, device=dev, dtype=torch.bfloat16)
    # per-batch modulation rows (row_div = T)
    ops.rmsnorm_modulate(x, out, 1e-6, mod, D, 0, table[1], table[0], row_div=T)
    xr = x * torch.rsqrt((x * x).mean(-1, keepdim=True) + 1e-6)
    rowb = torch.arange(R, device=dev) // T
    ref = xr * (1 + table[1] + mod[rowb, D:2 * D]) + table[0] + mod[rowb, 0:D]
    check("rmsnorm_modulate_div", out, ref, 4e-3)
    modt = torch.randn(R, 6 * D, device=dev, generator=g) * 0.5
    ops.rmsnorm_modulate(x, out, 1e-6, modt, 4 * D, 3 * D, table[4], table[3], row_div=1)
    ref = xr * (1 + table[4] + modt[:, 4 * D:5 * D]) + table[3] + modt[:, 3 * D:4 * D]
    check("rmsnorm_modulate_tok", out, ref, 4e-3)
    ops.rmsnorm_modulate(x, out, 1e-6)
    check("rmsnorm_plain", out, xr, 4e-3)
    idx = torch.randint(0, 5, (R,), device=dev, generator=g, dtype=torch.int32)
    modi = torch.randn(5, 6 * D, device=dev, generator=g)
    ops.rmsnorm_modulate(x, out, 1e-6, modi, D, 0, table[1], table[0], row_index=idx)
    ref = xr * (1 + table[1] + modi[idx.long(), D:2 * D]) + table[0] + modi[idx.long(), 0:D]
    check("rmsnorm_modulate_idx", out, ref, 4e-3)
    for Dx in (2048, 512, 1024 + 8):
        xs = torch.randn(77, Dx, device=dev, generator=g)
        os_ = torch.empty(77, Dx, device=dev, dtype=torch.bfloat16)
        ops.rmsnorm_modulate(xs, os_, 1e-6)
        check("rmsnorm_plain_D", os_, xs * torch.rsqrt((xs * xs).mean(-1, keepdim=True) + 1e-6), 4e-3, D=Dx)
    # layernorm modulate
    emb = torch.randn(R, D, device=dev, generator=g)
    t2 = torch.randn(2, D, device=dev, generator=g) * 0.1
    ops.layernorm_modulate(x, out, 1e-6, emb, t2[1], t2[0])
    ln = torch.nn.functional.layer_norm(x, (D,), eps=1e-6)
    ref = ln * (1 + t2[1] + emb) + t2[0] + emb
    check("layernorm_modulate", out, ref, 4e-3)
    # gate residual
    y = torch.randn(R, D, device=dev, generator=g).bfloat16()
    x2 = x.clone()
    ops.gate_residual(x2, y, mod, 2 * D, table[2], row_div=T)
    ref = x + y.float() * (table[2] + mod[rowb, 2 * D:3 * D])
    check("gate_residual", x2, ref, 1e-6)
    x2 = x.clone()
    ops.gate_residual(x2, y)
    check("gate_residual_nogate", x2, x + y.float(), 1e-6)
    # qk norm + rope
    for (H, dh) in ((32, 128), (32, 64), (4, 128), (4, 64)):
        Dq = H * dh
        qkv = torch.randn(R, 3 * Dq, device=dev, generator=g).bfloat16()
        q0 = qkv[:, Dq:2 * Dq].clone()
        wq = 1 + 0.1 * torch.randn(Dq, device=dev, generator=g)
        ang = torch.rand(1, H, T, dh // 2, device=dev, generator=g) * 6.28
        cos, sin = torch.cos(ang).contiguous(), torch.sin(ang).contiguous()
        ops.qknorm_rope(qkv[:, Dq:2 * Dq], B, T, H, dh, wq, 1e-6, cos, sin)
        qf = q0.float()
        qn = qf * torch.rsqrt((qf * qf).mean(-1, keepdim=True) + 1e-6) * wq
        qh = qn.reshape(B, T, H, 2, dh // 2)
        c, s = cos[0].transpose(0, 1)[None], sin[0].transpose(0, 1)[None]  # (1,T,H,dh/2)
        o1 = qh[..., 0, :] * c - s * qh[..., 1, :]
        o2 = qh[..., 1, :] * c + s * qh[..., 0, :]
        ref = torch.stack([o1, o2], dim=-2).reshape(R, Dq)
        check("qknorm_rope", qkv[:, Dq:2 * Dq], ref, 4e-3, H=H, dh=dh)
        q1 = q0.clone()
        ops.qknorm_rope(q1, B, T, H, dh, wq, 1e-6)
        check("qknorm_norope", q1, qn, 4e-3, H=H, dh=dh)
    # timestep embedding
    t = torch.tensor([1.0, 0.725, 0.05, 0.0, 0.99375], device=dev)
    te = torch.empty(5, 256, device=dev, dtype=torch.bfloat16)
    ops.timestep_embed(t, 1000.0, 256, te)
    ex = torch.exp(-math.log(10000.0) * torch.arange(128, device=dev, dtype=torch.float32) / 128)
    arg = (t * 1000.0)[:, None] * ex[None]
    check("timestep_embed", te, torch.cat([torch.cos(arg), torch.sin(arg)], -1), 4e-3)
    # silu / casts
    xb = torch.randn(1000 * 8 + 3, device=dev, generator=g).bfloat16()
    ob = torch.empty_like(xb)
    ops.silu_bf16(xb, ob)
    check("silu", ob, torch.nn.functional.silu(xb.float()), 4e-3)
    xf = torch.randn(4099, device=dev, generator=g)
    ob = torch.empty(4099, device=dev, dtype=torch.bfloat16)
    ops.cast_f32_to_bf16(xf, ob)
    check("cast_f32_bf16", ob, xf.bfloat16(), 0.0)
    of = torch.empty(4099, device=dev)
    ops.cast_bf16_to_f32(ob, of)
    check("cast_bf16_f32", of, ob.float(), 0.0)
    # euler step
    n, Cc = 640, 128
    lat = torch.randn(n, Cc, device=dev, generator=g)
    vp, vn = torch.randn(n, Cc, device=dev, generator=g), torch.randn(n, Cc, device=dev, generator=g)
    msk = (torch.rand(n, device=dev, generator=g) > 0.3).float()
    clean = torch.randn(n, Cc, device=dev, generator=g)
    sig, sign = 0.725, 0.421875
    st = (sig * msk).contiguous()
    lat2 = lat.clone()
    x0o = torch.empty_like(lat)
    ops.euler_step(lat2, vp, sig, sign, v_neg=vn, cfg_scale=4.5, sigma_tok=st, mask=msk, clean=clean, x0_out=x0o)
    v = vp + 3.5 * (vp - vn)
    x0 = lat - st[:, None] * v
    x0 = x0 * msk[:, None] + clean * (1 - msk[:, None])
    check("euler_x0", x0o, x0, 1e-6)
    check("euler_step", lat2, x0 + sign * (lat - x0) / sig, 1e-6)
    # bandwidth of the row kernel at the headline shape
    T2 = 5184
    xx = torch.randn(T2, D, device=dev)
    oo = torch.empty(T2, D, device=dev, dtype=torch.bfloat16)
    mm = torch.randn(1, 6 * D, device=dev)
    med, best = time_fn(lambda: ops.rmsnorm_modulate(xx, oo, 1e-6, mm, D, 0, table[1], table[0], row_div=T2), iters=20)
    emit(perf="rmsnorm_modulate", rows=T2, D=D, ms=med, gbs=T2 * D * 6 / med / 1e6)


GROUPS = {k[len("group_"):]: v for k, v in list(globals().items()) if k.startswith("group_")}

if __name__ == "__main__":
    name = sys.argv[1]
    t0 = time.time()
    emit(group=name, device=torch.cuda.get_device_name(0))
    GROUPS[name]()
    emit(group=name, done=True, failures=FAIL, seconds=round(time.time() - t0, 1))
    sys.exit(1 if FAIL else 0)
