"""Few-row GEMMs (sequence-parallel shards, audio tokens, AdaLN rows): microseconds per launch of the weight-streaming
kernel (cta_pair=4, gemm_small_m.cu) with explicit k-range split counts, of whatever the library picks on its own ("auto";
run once more with LTXB_GEMM_SMALL_M=0 for the big-tile kernel), and of cuBLAS.  Graph-replayed back-to-back launches, the
weights rotated over 8 copies so that they stream from HBM.  HBM floor = N * K * 2 bytes / 6551 GB/s."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
import mlx_video_b200  # noqa: E402,F401
from mlx_video_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
REPS, COPIES = 24, 8


def timed(fn):
    for r in range(COPIES):
        fn(r)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for r in range(REPS):
            fn(r % COPIES)
    g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / REPS * 1e3)
    return round(best, 1)


shapes = [(M, N, K) for M in (16, 68, 160, 256, 320) for N, K in ((4096, 4096), (12288, 4096), (16384, 4096), (4096, 16384), (2048, 2048))]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in s.split("x")) for s in sys.argv[1:]]
for M, N, K in shapes:
    a = torch.randn(M, K, device=dev).bfloat16()
    ws = [(torch.randn(N, K, device=dev) / 64).bfloat16() for _ in range(COPIES)]
    bias = torch.zeros(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    row = {"M": M, "N": N, "K": K, "hbm_floor_us": round(N * K * 2 / 6551e9 * 1e6, 1)}
    variants = [("auto", {}), ("small_m", dict(cta_pair=4))] + [(f"small_m_s{s}", dict(cta_pair=4, block_n=s)) for s in (1, 2, 4, 8)]
    only = os.environ.get("LTXB_BENCH_VARIANTS")
    if only:
        variants = [v for v in variants if v[0] in only.split(",")]
    for name, kw in variants:
        try:
            row[name] = timed(lambda r: ops.gemm(a, ws[r], bias, out, mode=_lib.EPI_BIAS_BF16, **kw))
        except Exception as e:  # noqa: BLE001
            row[name] = str(e)[:60]
    if os.environ.get("LTXB_BENCH_PACKED", "1") != "0":
        for bits in (8, 4):  # packed weights (ltxb_gemm_qw_bf16): random words, unit scales — timing only
            pk = [torch.randint(-2**31, 2**31 - 1, (N, K * bits // 32), device=dev, dtype=torch.int32) for _ in range(COPIES)]
            sc = torch.full((N, K // 64), 0.01, device=dev, dtype=torch.bfloat16)
            bi = torch.zeros(N, K // 64, device=dev, dtype=torch.bfloat16)
            try:
                row[f"packed_w{bits}"] = timed(lambda r: ops.gemm_qw(a, pk[r], sc, bi, 64, bits, bias, out, mode=_lib.EPI_BIAS_BF16, const_w=True))
            except Exception as e:  # noqa: BLE001
                row[f"packed_w{bits}"] = str(e)[:60]
            del pk
    row["cublas"] = timed(lambda r: torch.matmul(a, ws[r].T, out=out))
    print(json.dumps(row), flush=True)
