"""Print the phase timeline of CTA 0 of the pair attention kernel (library built with -DLTXB_ATTN_TRACE)."""
import math
import sys

import torch

sys.path.insert(0, ".")
import mlx_video_b200  # noqa: E402,F401
from mlx_video_b200 import ops  # noqa: E402

B, Tq, Tk, H, dh = (int(v) for v in sys.argv[1:6]) if len(sys.argv) > 5 else (1, 5184, 5184, 32, 128)
dev = torch.device("cuda:0")
D = H * dh
q, k, v = (torch.randn(B * T, D, device=dev).bfloat16() for T in (Tq, Tk, Tk))
out = torch.empty(B * Tq, D, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    ops.attention(q, k, v, out, B, Tq, Tk, H, dh, 1 / math.sqrt(dh))
torch.cuda.synchronize()
tr = ops._attn_workspaces[0][:4 * 16 * 8 * 8].view(torch.int64).view(4, 16, 8).cpu()
t0 = int(tr[3, 0, 0])  # kernel entry of CTA 0
names = [["wait", "S rdy", "ld ok", "max ok", "P lo", "P hi", "-", "-"]] * 2 + [["P0lo", "P0hi", "PV0 is", "S0 is", "P1lo", "P1hi", "PV1 is", "S1 is"]]
for role, rn in enumerate(["softmax0", "softmax1", "mma"]):
    print(rn, names[role])
    for it in range(16):
        print(f"  it {it:2d}: " + " ".join(f"{int(x) - t0:7d}" if int(x) else "      -" for x in tr[role, it]))
print("cta life cycle [entry, setup done, predecessor done, softmax0 done, softmax1 done, last PV done, all stored]:")
print("  " + " ".join(f"{int(x) - t0:7d}" for x in tr[3, 0, :7]))
