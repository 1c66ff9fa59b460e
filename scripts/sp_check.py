"""Multi-GPU parity check (run under torchrun, one rank per GPU): sequence-parallel and CFG-parallel results
against the single-GPU model on the same weights/inputs, and against the fp32 oracle."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
sys.path.insert(0, "oracle")
import ltx_oracle as O  # noqa: E402
import mlx_video_b200 as M  # noqa: E402
from mlx_video_b200 import parallel, sampler  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
fails = []


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def report(name, err, tol):
    ok = err <= tol
    if rank == 0:
        print(f"{'ok  ' if ok else 'FAIL'} {name}: rel_l2 {err:.3e} (tol {tol:g})", flush=True)
    if not ok:
        fails.append(name)


def product_config(cfg):
    d = {k: getattr(cfg, k) for k in cfg.__dataclass_fields__}
    d["model_type"], d["rope_type"] = d["model_type"].value, d["rope_type"].value
    return M.LTXModelConfig.from_dict(d)


def to_dev(m):
    return None if m is None else M.Modality(m.latent.to(dev), m.timesteps.to(dev), m.positions.to(dev), m.context.to(dev))


heads = max(4, world)
for mt, name in [(O.LTXModelType.VideoOnly, "video"), (O.LTXModelType.AudioVideo, "audio+video")]:
    cfg = O.small_config(mt, num_layers=2, heads=heads, audio_heads=heads)
    tensors = O.init_params(cfg, seed=3)
    tensors = {k: (v.to(torch.bfloat16).float() if k.endswith(".weight") and not k.endswith("_norm.weight") else v) for k, v in tensors.items()}
    g = torch.Generator().manual_seed(4)
    F_, H_, W_ = 4, 4, 2 * world
    T, Tc, Ta = F_ * H_ * W_, 24, 21
    ts = torch.full((1, T), 0.725)
    ts[:, :H_ * W_] = 0.0
    video = O.Modality(torch.randn(1, T, 128, generator=g), ts, torch.from_numpy(O.create_position_grid(1, F_, H_, W_)), torch.randn(1, Tc, 256, generator=g))
    audio = None
    if mt == O.LTXModelType.AudioVideo:
        audio = O.Modality(torch.randn(1, Ta, 128, generator=g), torch.full((1, Ta), 0.725), torch.from_numpy(O.create_audio_position_grid(1, Ta)), torch.randn(1, Tc, 256, generator=g))
    wv, wa = O.OracleLTXModel(cfg, tensors)(video, audio)
    model = M.LTXModel(product_config(cfg), device=dev)
    model.load_weights(tensors)
    sv, sa = model(video=to_dev(video), audio=to_dev(audio))  # single GPU
    for fused in (False, True):
        layout = parallel.make_layout(world, rank, use_cfg=False, fused=fused)
        layout.attach(model)
        for rep in range(2):  # twice: the second pass reuses the peer buffers / flag epochs
            pv, pa = model(video=to_dev(video), audio=to_dev(audio))  # sequence parallel over all ranks
        model.seq_parallel = None
        tag = f"{name}: {layout.describe()}"
        report(f"{tag} vs oracle (video)", rel(pv, wv), 1e-2)
        report(f"{tag} vs single GPU (video)", rel(pv, sv), 5e-3)
        if audio is not None:
            report(f"{tag} vs oracle (audio)", rel(pa, wa), 1e-2)
            report(f"{tag} vs single GPU (audio)", rel(pa, sa), 5e-3)

# CFG-parallel (x Ulysses when world > 2) dev sampler vs the single-GPU cfg_batch sampler
if world % 2 == 0:
    cfg = O.small_config(O.LTXModelType.VideoOnly, num_layers=2, heads=heads)
    tensors = O.init_params(cfg, seed=5)
    model = M.LTXModel(product_config(cfg), device=dev)
    model.load_weights(tensors)
    g = torch.Generator().manual_seed(6)
    b, c, f, h, w = 1, 128, 3, 4, 2 * world
    lat = torch.randn(b, c, f, h, w, generator=g).to(dev)
    pos = sampler.create_position_grid(b, f, h, w)
    cp, cn = torch.randn(b, 16, 256, generator=g).to(dev), torch.randn(b, 16, 256, generator=g).to(dev)
    sig = sampler.ltx2_scheduler(3, f * h * w)
    single = sampler.denoise_dev(lat, pos, cp, cn, model, sig, cfg_scale=4.5, cfg_batch=True)
    layout = parallel.make_layout(world, rank, use_cfg=True)
    layout.attach(model)
    par = sampler.denoise_dev(lat, pos, cp, cn, model, sig, cfg_scale=4.5, cfg_parallel=layout.cfg)
    report(f"denoise_dev {layout.describe()} vs single-GPU cfg_batch", rel(par, single), 5e-3)
    gathered = [torch.empty_like(par) for _ in range(world)]
    dist.all_gather(gathered, par.contiguous())
    report("latents identical on every rank", max(rel(t, gathered[0]) for t in gathered), 0.0)

flag = torch.tensor([len(fails)], device=dev)
dist.all_reduce(flag)
dist.destroy_process_group()
sys.exit(1 if int(flag) else 0)
