"""LTX-2 video VAE decoder at a production size (768x768x65: latents 128 x 9 x 24 x 24): ms per decode, convolution TFLOP/s.
Random-init full-width weights (545 M parameters).  Usage: python scripts/vae_bench.py [F H W] [--tiled]"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
import mlx_video_b200  # noqa: E402,F401
from mlx_video_b200 import ops  # noqa: E402
from mlx_video_b200 import vae_decoder as VD  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
F_, H, W = (int(v) for v in args[:3]) if len(args) >= 3 else (9, 24, 24)
dev = torch.device("cuda:0")
model = VD.LTX2VideoDecoder(device=dev)
g = torch.Generator(device=dev).manual_seed(0)
for name, p in model.parameters().items():
    if name == "timestep_scale_multiplier":
        continue
    if name.endswith(".weight"):
        p.copy_((torch.rand(p.shape, generator=g, device=dev) * 2 - 1) / (p.shape[1] ** 0.5))
    elif name.endswith("table"):
        p.copy_(0.05 * torch.randn(p.shape, generator=g, device=dev))
    elif name == "latents_std":
        p.fill_(1.0)
x = torch.randn(1, 128, F_, H, W, generator=g, device=dev)


def conv_flops(F_, H, W):
    total, pos = 0.0, F_ * H * W
    widths = (1024, 512, 256, 128)
    total += pos * 54.0 * 128 * 1024
    for lvl, c in enumerate(widths):
        total += pos * 54.0 * 10 * c * c
        if lvl < 3:
            total += pos * 54.0 * c * 4 * c
            pos = (2 * (pos // (H * W * 4 ** lvl)) - 1) * (H * W * 4 ** (lvl + 1))
    total += pos * 54.0 * 128 * 48
    return total


def run():
    if "--tiled" in sys.argv:
        return model.decode_tiled(x, VD.TilingConfig.default())
    return model(x)


l0 = ops.launches
out = run()
torch.cuda.synchronize()
launches = ops.launches - l0
ts = []
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = run()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[1]
fl = conv_flops(F_, H, W)
print(json.dumps({"latents": [1, 128, F_, H, W], "video": list(out.shape), "ms_per_decode": round(ms, 2), "conv_tflop": round(fl / 1e12, 2),
                  "conv_tflops_per_s": round(fl / ms / 1e9), "kernel_launches": launches, "tiled": "--tiled" in sys.argv,
                  "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}))
