"""Latent upsampler at production width (mid 1024, 4 + 4 res blocks) on the stage-1 latent of BASELINE configs[1]
(128 x 5 x 16 x 16 -> 128 x 5 x 32 x 32): device time per call (CUDA events), conv FLOP rate, and the oracle on the host."""
import json, sys, time
import torch
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import mlx_video_b200 as M

dev = torch.device("cuda:0")
F_, H, W, mid = 5, 16, 16, 1024
model = M.LatentUpsampler(128, mid, 4, device=dev)
g = torch.Generator(device=dev).manual_seed(0)
for _, conv in model._convs():
    conv.weight.copy_((torch.rand(conv.weight.shape, generator=g, device=dev) * 2 - 1) / conv.weight.shape[1] ** 0.5)
lat = torch.randn(1, 128, F_, H, W, device=dev, generator=g)
mean, std = torch.zeros(128, device=dev), torch.ones(128, device=dev)
S = F_ * H * W
flops = 2.0 * 27 * (S * 128 * mid + 8 * S * mid * mid + 8 * 4 * S * mid * mid + 4 * S * mid * 128) + 2.0 * 9 * S * mid * 4 * mid
for _ in range(3):
    out = M.upsample_latents(lat, model, mean, std)
torch.cuda.synchronize()
n0 = M.ops.launches
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    out = M.upsample_latents(lat, model, mean, std)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
res = dict(workload="latent upsampler 128x5x16x16 -> 128x5x32x32, mid 1024, 4+4 res blocks", ms_per_call=ms, conv_tflop=flops / 1e12,
           tflops=flops / ms / 1e9, launches_per_call=(M.ops.launches - n0) // 5, finite=bool(torch.isfinite(out).all()))
if "--cpu" in sys.argv:
    import upsampler_oracle as U
    p = U.init_upsampler_params(128, mid, 4, seed=0)
    t0 = time.perf_counter(); U.upsample_latents(lat.cpu(), p, mean.cpu(), std.cpu()); res["oracle_cpu_s"] = time.perf_counter() - t0
    res["cpu_threads"] = torch.get_num_threads()
print(json.dumps(res))
