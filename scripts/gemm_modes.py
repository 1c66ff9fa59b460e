"""Epilogue cost of the GEMM at the LTX-2 N = 4096 shapes: same problem, auto schedule, different epilogue modes.
L2 is flushed between launches (a 512 MB memset), every launch timed with CUDA events; prints the median.
args: [iters] [ncu]   (with `ncu`: two launches per variant only, no flush timing — for a profiler run)"""
import json
import math
import sys

import torch

sys.path.insert(0, ".")
import mlx_video_b200  # noqa: E402,F401
from mlx_video_b200 import _lib, ops  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def run(M, N, K, mode, gate, inplace, tag):
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device=dev)
    out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16 if mode < 3 else torch.float32)
    kw = {}
    if mode == _lib.EPI_RESID_GATE_F32:
        kw["resid"] = out if inplace else torch.zeros_like(out)
        if gate:
            kw.update(gate=torch.randn(1, N, device=dev), gate_table=torch.randn(N, device=dev), gate_row_div=M)
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gemm(a, w, bias, out, mode=mode, **kw)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    print(json.dumps(dict(tag=tag, M=M, N=N, K=K, us=round(ms * 1e3, 1), tflops=round(2.0 * M * N * K / ms / 1e9))), flush=True)


for K in (4096, 16384):
    run(1280, 4096, K, _lib.EPI_BIAS_BF16, False, False, "bias->bf16")
    run(1280, 4096, K, _lib.EPI_BIAS_F32, False, False, "bias->f32")
    run(1280, 4096, K, _lib.EPI_RESID_GATE_F32, False, False, "resid (out of place)")
    run(1280, 4096, K, _lib.EPI_RESID_GATE_F32, False, True, "resid in place")
    run(1280, 4096, K, _lib.EPI_RESID_GATE_F32, True, True, "resid in place + gate")
