"""Run a few launches of one GEMM configuration (for ncu / quick timing). args: M N K pair bn [iters] [mode]"""
import math, sys, json
import torch
sys.path.insert(0, ".")
import mlx_video_b200  # noqa
from mlx_video_b200 import ops
M, N, K, pair, bn = (int(v) for v in sys.argv[1:6])
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 5
mode = int(sys.argv[7]) if len(sys.argv) > 7 else 0
dev = torch.device("cuda:0")
a = torch.randn(M, K, device=dev).bfloat16()
w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
bias = torch.randn(N, device=dev)
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16 if mode < 3 else torch.float32)
kw = {}
if mode == 4:
    kw = dict(resid=out)
    out.zero_()
ts = []
for i in range(iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.gemm(a, w, bias, out, mode=mode, block_n=bn, cta_pair=pair, **kw); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ts.sort()
print(json.dumps(dict(M=M, N=N, K=K, pair=pair, bn=bn, mode=mode, ms=ts[len(ts)//2], tflops=2.0*M*N*K/ts[len(ts)//2]/1e9)))
