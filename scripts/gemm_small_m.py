"""Small-M GEMMs (sequence-parallel shards: M = 160 / 320 / 640 rows against the full LTX-2 weights), L2 flushed between
runs: microseconds and streamed-weight GB/s per schedule.  Usage: python scripts/gemm_small_m.py [M ...]"""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
import mlx_video_b200  # noqa: E402,F401
from mlx_video_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)


def time_fn(fn, iters=15, warmup=3):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        torch.cuda._sleep(400000)  # ~0.2 ms of GPU idle-spin: the host gets ahead, launch latency stays out of the bracket
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3  # us


Ms = [int(v) for v in sys.argv[1:]] or [160, 320, 640]
for M in Ms:
    for N, K in [(12288, 4096), (4096, 4096), (16384, 4096), (4096, 16384)]:
        a = torch.randn(M, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) / 64).bfloat16()
        bias = torch.zeros(N, device=dev)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        wbytes = 2.0 * N * K
        row = {"M": M, "N": N, "K": K, "hbm_floor_us": round(wbytes / 6551e3, 1)}
        for name, kw in [("cublas", None), ("auto", {}), ("lockstep_pair256", dict(block_n=256, cta_pair=1)), ("contig_single", dict(cta_pair=2)),
                         ("contig_pair", dict(cta_pair=3)), ("pair224", dict(block_n=224, cta_pair=1)), ("pair176", dict(block_n=176, cta_pair=1)), ("single128", dict(block_n=128, cta_pair=0))]:
            try:
                us = time_fn((lambda: torch.matmul(a, w.T, out=out)) if kw is None else (lambda: ops.gemm(a, w, bias, out, **kw)))
                row[name] = {"us": round(us, 1), "w_gbs": round(wbytes / us / 1e3)}
            except Exception as e:  # noqa: BLE001
                row[name] = str(e)[:60]
        print(json.dumps(row), flush=True)
