"""Where a small GEMM's time goes: duration against K (fixed cost vs per-k-block cost), graph-replayed back-to-back launches
(no host launch latency, weights L2-cold by rotation over 8 weight copies)."""
import json
import sys

import torch

sys.path.insert(0, ".")
import mlx_video_b200  # noqa: E402,F401
from mlx_video_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def bench(M, N, K, reps=24, copies=8, **kw):
    a = torch.randn(M, K, device=dev).bfloat16()
    ws = [(torch.randn(N, K, device=dev) / 64).bfloat16() for _ in range(copies)]
    bias = torch.zeros(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for w in ws:
        ops.gemm(a, w, bias, out, **kw)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for r in range(reps):
            ops.gemm(a, ws[r % copies], bias, out, **kw)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for M in (160, 1280):
    for N in (4096, 16384):
        for K in (64, 512, 2048, 4096, 16384):
            if N * K > 16384 * 4096:
                continue
            row = {"M": M, "N": N, "K": K}
            for name, kw in [("auto", {}), ("pair256_dp", dict(block_n=256, cta_pair=1)), ("single256_dp", dict(block_n=256, cta_pair=0)), ("contig_pair", dict(cta_pair=3))]:
                try:
                    row[name] = round(bench(M, N, K, **kw), 1)
                except Exception as e:  # noqa: BLE001
                    row[name] = str(e)[:40]
            a = torch.randn(M, K, device=dev).bfloat16()
            ws = [(torch.randn(N, K, device=dev) / 64).bfloat16() for _ in range(8)]
            out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            g = torch.cuda.CUDAGraph()
            torch.matmul(a, ws[0].T, out=out)
            torch.cuda.synchronize()
            with torch.cuda.graph(g):
                for r in range(24):
                    torch.matmul(a, ws[r % 8].T, out=out)
            g.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            row["cublas"] = round(e0.elapsed_time(e1) / 24 * 1e3, 1)
            print(json.dumps(row), flush=True)
