"""Per-CTA life cycle of the small-M GEMM kernel (debug build: LTXB_LIB=libltxb_wsdbg.so LTXB_WS_DEBUG=8): nanoseconds from
the earliest PDL-wait exit to accumulator complete / split partners met / epilogue done, over all CTAs of one launch."""
import sys

import torch

sys.path.insert(0, ".")
import mlx_video_b200  # noqa: E402,F401
from mlx_video_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
M, N, K = (int(v) for v in sys.argv[1].split("x"))
splits = int(sys.argv[2]) if len(sys.argv) > 2 else 0
a = torch.randn(M, K, device=dev).bfloat16()
ws = [(torch.randn(N, K, device=dev) / 64).bfloat16() for _ in range(8)]
bias = torch.zeros(N, device=dev)
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
ops.gemm(a, ws[0], bias, out, mode=_lib.EPI_BIAS_BF16, cta_pair=4, block_n=splits)
torch.cuda.synchronize()
ops._gemm_workspaces[0][-296 * 4 * 4 * 8:].zero_()
for r in range(16):
    ops.gemm(a, ws[r % 8], bias, out, mode=_lib.EPI_BIAS_BF16, cta_pair=4, block_n=splits)
torch.cuda.synchronize()
buf = ops._gemm_workspaces[0]
t = buf[-296 * 4 * 4 * 8:].view(torch.int64).view(-1, 4)[:296].cpu()
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
rel = (t - t0).float()
names = ["pdl_exit", "accumulator", "met", "done"]
print(f"{M}x{N}x{K} splits={splits}: {t.shape[0]} CTAs, kernel span {float(rel[:, 3].max()) / 1e3:.1f} us")
for i, n in enumerate(names):
    c = rel[:, i]
    print(f"  {n:12s} min {c.min() / 1e3:6.1f}  median {c.median() / 1e3:6.1f}  max {c.max() / 1e3:6.1f} us")
d = (t[:, 2] - t[:, 1]).float()
print(f"  accumulator -> met: median {d.median() / 1e3:.1f}, max {d.max() / 1e3:.1f} us;  met -> done: median {((t[:, 3] - t[:, 2]).float().median()) / 1e3:.1f} us, max {((t[:, 3] - t[:, 2]).float().max()) / 1e3:.1f}")
