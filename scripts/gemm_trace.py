"""Life cycle of CTA 0 of one GEMM launch (library built with -DLTXB_GEMM_TRACE), in cycles from kernel entry; the launch is
preceded by another GEMM so the programmatic-launch chain looks like the model's.  args: M N K [pair bn]"""
import sys

import torch

sys.path.insert(0, ".")
import mlx_video_b200  # noqa: E402,F401
from mlx_video_b200 import ops  # noqa: E402

M, N, K = (int(v) for v in sys.argv[1:4])
pair = int(sys.argv[4]) if len(sys.argv) > 4 else -1
bn = int(sys.argv[5]) if len(sys.argv) > 5 else 0
dev = torch.device("cuda:0")
a = torch.randn(M, K, device=dev).bfloat16()
ws = [(torch.randn(N, K, device=dev) / 64).bfloat16() for _ in range(4)]
bias = torch.zeros(N, device=dev)
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
names = ["entry", "setup done", "predecessor done", "first operands", "MMAs issued", "accumulator full", "epilogue done", "teardown sync", "TMEM released", "owner: partials parked", "owner: first chunk done"]
for rep in range(3):
    for w in ws:
        ops.gemm(a, w, bias, out, cta_pair=pair, block_n=bn)
    torch.cuda.synchronize()
    buf = ops._gemm_workspaces[0]
    tr = buf[(1 << 16) * 4:(1 << 16) * 4 + 16 * 8].view(torch.int64).cpu().tolist()
    t0 = tr[0]
    if rep == 2:
        print(f"M={M} N={N} K={K} pair={pair} bn={bn}: " + ", ".join(f"{n} {tr[i] - t0}" for i, n in enumerate(names)))
