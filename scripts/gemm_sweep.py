"""GEMM timing at the LTX-2 shapes (L2 flushed between runs) vs torch.matmul (cuBLAS) as the library yardstick."""
import json
import sys

import torch

sys.path.insert(0, ".")
import mlx_video_b200  # noqa: E402,F401
from mlx_video_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)


def time_fn(fn, iters=15, warmup=3):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


shapes = [(1280, 12288, 4096), (1280, 4096, 4096), (1024, 8192, 4096), (1280, 16384, 4096), (1280, 4096, 16384),
          (5184, 12288, 4096), (5184, 4096, 4096), (5184, 16384, 4096), (5184, 4096, 16384), (2560, 12288, 4096), (2560, 4096, 16384),
          (14080, 16384, 4096), (648, 12288, 4096), (648, 4096, 16384)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in s.split("x")) for s in sys.argv[1:]]
for M, N, K in shapes:
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) / 64).bfloat16()
    bias = torch.zeros(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    fl = 2.0 * M * N * K
    row = {"M": M, "N": N, "K": K}
    row["cublas"] = round(fl / time_fn(lambda: torch.matmul(a, w.T, out=out)) / 1e9)
    row["auto"] = round(fl / time_fn(lambda: ops.gemm(a, w, bias, out)) / 1e9)
    row["dp_pair256"] = round(fl / time_fn(lambda: ops.gemm(a, w, bias, out, block_n=256, cta_pair=1)) / 1e9)
    row["dp_single256"] = round(fl / time_fn(lambda: ops.gemm(a, w, bias, out, block_n=256, cta_pair=0)) / 1e9)
    print(json.dumps(row), flush=True)
