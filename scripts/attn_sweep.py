"""Attention timing at the LTX-2 shapes vs torch SDPA (library yardstick)."""
import json
import math
import os
import sys

import torch

sys.path.insert(0, ".")
import mlx_video_b200  # noqa: E402,F401
from mlx_video_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def time_fn(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for (B, T, Tk, H, dh) in [(1, 1280, 1280, 32, 128), (1, 1280, 1024, 32, 128), (2, 5184, 5184, 32, 128), (2, 5184, 1024, 32, 128), (1, 14080, 14080, 32, 128), (1, 5184, 68, 32, 64)]:
    q = torch.randn(B * T, H * dh, device=dev).bfloat16()
    k = torch.randn(B * Tk, H * dh, device=dev).bfloat16()
    v = torch.randn(B * Tk, H * dh, device=dev).bfloat16()
    o = torch.empty_like(q)
    fl = 4.0 * B * H * T * Tk * dh
    ms = time_fn(lambda: ops.attention(q, k, v, o, B, T, Tk, H, dh, 1 / math.sqrt(dh)))
    q4, k4, v4 = (t.view(B, -1, H, dh).transpose(1, 2) for t in (q, k, v))
    ms_t = float("inf") if os.environ.get("ATTN_SWEEP_NO_SDPA") else time_fn(lambda: torch.nn.functional.scaled_dot_product_attention(q4, k4, v4))
    print(json.dumps({"B": B, "T": T, "Tk": Tk, "H": H, "dh": dh, "ltxb_us": round(ms * 1e3, 1), "ltxb_tflops": round(fl / ms / 1e9), "sdpa_tflops": round(fl / ms_t / 1e9)}), flush=True)
