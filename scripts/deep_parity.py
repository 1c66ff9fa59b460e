"""Production-width depth parity: L blocks of the real LTX-2 width (D=4096, 32x128 heads) on the GPU vs the fp32
oracle on the host, same (bf16-rounded) weights and inputs.  BASELINE bar: rel-L2 <= 1e-2, cosine >= 0.999.
Usage: python scripts/deep_parity.py [L] [T] [Tc]      (host RAM: ~1.1 GB of fp32 weights per block)"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "oracle")
import ltx_oracle as O  # noqa: E402
import mlx_video_b200 as M  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 320
Tc = int(sys.argv[3]) if len(sys.argv) > 3 else 128
grid = {320: (5, 8, 8), 1280: (5, 16, 16)}[T]
torch.set_num_threads(torch.get_num_threads())
cfg = O.OracleConfig(num_layers=L)
t0 = time.time()
tensors = O.init_params(cfg, seed=0)
for k, v in tensors.items():
    if k.endswith(".weight") and not k.endswith("_norm.weight"):
        tensors[k] = v.to(torch.bfloat16).float()
g = torch.Generator().manual_seed(1)
video = O.Modality(torch.randn(1, T, 128, generator=g), torch.full((1, T), 0.725), torch.from_numpy(O.create_position_grid(1, *grid)),
                   torch.randn(1, Tc, 3840, generator=g))
model = M.LTXModel(M.production_config(M.LTXModelType.VideoOnly, num_layers=L), device="cuda:0")
model.load_weights(tensors)
got, _ = model(video=M.Modality(video.latent.cuda(), video.timesteps.cuda(), video.positions.cuda(), video.context.cuda()), audio=None)
torch.cuda.synchronize()
t1 = time.time()
with torch.no_grad():
    oracle = O.OracleLTXModel(cfg, tensors)
    va, _ = oracle.prepare(video, None)
    per_block = []
    for i in range(L):
        va, _ = oracle.block(i, va, None)
    want = oracle._output(oracle.p["scale_shift_table"], oracle.p.sub("proj_out"), va.x, va.embedded_timestep)
rel = float((got.cpu().double() - want.double()).norm() / want.double().norm())
cos = float(torch.nn.functional.cosine_similarity(got.cpu().double().flatten(), want.double().flatten(), dim=0))
print(json.dumps({"layers": L, "T": T, "Tc": Tc, "rel_l2": rel, "cosine": cos, "ok": rel <= 1e-2 and cos >= 0.999,
                  "setup_s": round(t1 - t0, 1), "oracle_s": round(time.time() - t1, 1)}), flush=True)
