"""Run a few back-to-back launches of one attention shape (for ncu / quick timing). args: B Tq Tk H dh [iters]"""
import json
import math
import sys

import torch

sys.path.insert(0, ".")
import mlx_video_b200  # noqa: E402,F401
from mlx_video_b200 import ops  # noqa: E402

B, Tq, Tk, H, dh = (int(v) for v in sys.argv[1:6])
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 5
dev = torch.device("cuda:0")
D = H * dh
q = torch.randn(B * Tq, D, device=dev).bfloat16()
k = torch.randn(B * Tk, D, device=dev).bfloat16()
v = torch.randn(B * Tk, D, device=dev).bfloat16()
out = torch.empty(B * Tq, D, device=dev, dtype=torch.bfloat16)
scale = 1.0 / math.sqrt(dh)
for _ in range(2):
    ops.attention(q, k, v, out, B, Tq, Tk, H, dh, scale)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    ops.attention(q, k, v, out, B, Tq, Tk, H, dh, scale)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / iters * 1e3
print(json.dumps(dict(B=B, Tq=Tq, Tk=Tk, H=H, dh=dh, us=round(us, 1), tflops=round(4.0 * B * H * Tq * Tk * dh / us / 1e6))))
