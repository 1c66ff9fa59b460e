#!/usr/bin/env python
"""LTX-2 DiT denoise-step throughput on B200 (BASELINE.json metric: video tokens/s and steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload distilled|dev|av|long]
                    [--workloads dev,av,long|none]

A *step* is one denoising step of the sampler loop: build the Modality, run the 48-block DiT forward
(two forwards per step under CFG), fused CFG / x0 / Euler update of the latent.  Workloads:
  distilled  BASELINE configs[1]: full 48-block video-only LTX-2 DiT, 512x512x33 -> 16x16x5 = 1280 video
             tokens, 1024 text tokens (what the reference pipelines pass), STAGE_1_SIGMAS,
             no CFG, bf16 tensor-core math with fp32 residual stream.  N > 1: Ulysses sequence parallel
             (strong scaling: the same video on N GPUs).
  dev        BASELINE configs[2]: 768x768x65 -> 24x24x9 = 5184 tokens, CFG 4.5 (cond + uncond forward per
             step).  N = 1: cfg_batch (B=2).  N >= 2: CFG-parallel x Ulysses-(N/2).
  av         BASELINE configs[3]: the joint audio+video model (5184 video + 68 audio tokens).  N > 1: Ulysses-N.
  long       BASELINE configs[4]: 1280x704x121 -> 14080 video tokens, distilled stage 2.  N > 1: Ulysses-N.
ONE JSON line on stdout (rank 0): the headline workload (`--workload`, default distilled = the config the metric is
quoted on) with the full contract, plus `"workloads": {...}` — the other BASELINE configs measured for a few steps each
in the same run (same fields: tokens/s, ms/step, parallelism, roofline, parity), so the driver's BENCH / SCALE records
carry CFG-parallel x Ulysses, the joint audio+video model and the 14k-token config too.
`value` has inputs resident in HBM; `e2e` goes through the public API with pinned HOST buffers copied in and the
velocity read back every step.  The only places this file touches oracle/ are the CPU baseline (`cpu_baseline`,
`--impl reference`) and the small-model parity check that runs BEFORE the timed region (checker, never measured).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (frames, height, width, text tokens, cfg_scale, sigmas)
    "distilled": dict(grid=(5, 16, 16), Tc=1024, cfg=1.0, desc="LTX-2 19B video-only DiT, 48 blocks, distilled stage-1, 512x512x33 (16x16x5=1280 tokens), 1024 text tokens"),
    "dev": dict(grid=(9, 24, 24), Tc=1024, cfg=4.5, desc="LTX-2 19B video-only DiT, 48 blocks, dev pipeline, 768x768x65 (24x24x9=5184 tokens), CFG 4.5, 1024 text tokens"),
    # BASELINE configs[4]: distilled stage 2 of 1280x704x121 -> 40x22x16 = 14080 video tokens
    "long": dict(grid=(16, 22, 40), Tc=1024, cfg=1.0, desc="LTX-2 19B video-only DiT, 48 blocks, distilled stage-2, 1280x704x121 (40x22x16=14080 tokens), 1024 text tokens"),
    # BASELINE configs[3]: joint audio+video model, 768x768x65 + 68 audio latents (audio<->video cross-attention in every block)
    # debug stand-in (not a BASELINE config): the GEMM rows ONE rank sees in the 8-GPU distilled run (1280 / 8 = 160 tokens)
    "shard160": dict(grid=(1, 10, 16), Tc=1024, cfg=1.0, desc="debug: 160 video tokens on one GPU = the per-rank GEMM shapes of the 8-GPU distilled run"),
    "av": dict(grid=(9, 24, 24), Tc=1024, cfg=1.0, Ta=68, desc="LTX-2 19B audio+video DiT, 48 blocks, 768x768x65 (5184 video + 68 audio tokens), 1024 text tokens per modality"),
}


def peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(tflops=float(d.get("bf16_tflops_sustained", 1386.9)), burst=float(d.get("bf16_tflops", 1652.9)),
                    hbm=float(d.get("hbm_gbs", 6551.0)), source="measured")
    return dict(tflops=1400.0, burst=1590.0, hbm=6650.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int) -> None:
        self.index, self.proc, self.lines = index, None, []

    def start(self) -> None:
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.thread.join(timeout=2)
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])), smax.append(float(f[1])), power.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power), "samples": len(sm),
                "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference forward on the host cores (reported, not the target)
# --------------------------------------------------------------------------------------------------
class CpuSample:
    """Bounded sample of the workload on the CPU: ONE production-width block (of 48) plus the model's pre/post
    processing, fp32 oracle port, all host threads; a full forward is extrapolated as 48 x block + pre/post."""

    def __init__(self, T: int, Tc: int) -> None:
        sys.path.insert(0, str(ROOT / "oracle"))
        import torch

        import ltx_oracle as O

        self.torch, self.T, self.Tc = torch, T, Tc
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        cfg = O.OracleConfig(num_layers=1)
        self.model = O.OracleLTXModel(cfg, O.init_params(cfg, seed=0))
        g = torch.Generator().manual_seed(1)
        F_, H_, W_ = {1280: (5, 16, 16), 5184: (9, 24, 24), 14080: (16, 22, 40)}.get(T, (1, 1, T))
        video = O.Modality(torch.randn(1, T, 128, generator=g), torch.full((1, T), 0.725),
                           torch.from_numpy(O.create_position_grid(1, F_, H_, W_)), torch.randn(1, Tc, 3840, generator=g))
        with torch.no_grad():
            t0 = time.perf_counter()
            self.va, _ = self.model.prepare(video, None)
            m = self.model
            m._output(m.p["scale_shift_table"], m.p.sub("proj_out"), self.va.x, self.va.embedded_timestep)
            self.t_prepost = time.perf_counter() - t0

    def block(self) -> float:
        with self.torch.no_grad():
            t0 = time.perf_counter()
            self.model.block(0, self.va, None)
            return time.perf_counter() - t0

    def describe(self, times) -> str:
        return (f"1 of 48 blocks (fp32 oracle port, T={self.T}, Tc={self.Tc}, D=4096) x{len(times)} reps, median "
                f"{statistics.median(times) * 1e3:.0f} ms, + pre/post {self.t_prepost * 1e3:.0f} ms; forward = 48 x block + pre/post (extrapolated)")


def cpu_block_sample(T: int, Tc: int, budget_s: float = 15.0) -> dict:
    s = CpuSample(T, Tc)
    s.block()  # warm-up
    times, t_start = [], time.perf_counter()
    while len(times) < 2 or (time.perf_counter() - t_start < budget_s and len(times) < 8):
        times.append(s.block())
    return dict(t_forward=48 * statistics.median(times) + s.t_prepost, cores=s.cores, sample=s.describe(times))


def workload_config(wl: dict, layers: int, batch: int, forwards: int, parallelism: str, cache: bool) -> dict:
    """The `config` object of the JSON line — the SAME keys on both arms (ours / reference)."""
    F_, H_, W_ = wl["grid"]
    return {"workload": wl["desc"], "video_tokens": F_ * H_ * W_, "audio_tokens": wl.get("Ta", 0), "text_tokens": wl["Tc"],
            "layers": layers, "batch": batch, "forwards_per_step": forwards, "parallelism": parallelism,
            "l2": "48 blocks x 537 MB of bf16 weights stream through the 126 MB L2 every step (inputs larger than L2)",
            "weights": "random-init, seeded, bf16; fp32 residual stream",
            "text_kv": ("projected once per denoise loop and reused across its steps (the context is loop-invariant; row N1)" if cache
                        else "recomputed every step, as the reference does")}


def run_reference(args, wl, rank: int, result_out) -> None:
    """`--impl reference`: the reference's CPU implementation of the path.  The reference itself is MLX-only
    Python and cannot run here, so this is the oracle port (kind "port") on all host threads; every
    step is a bounded sample (one block of 48), extrapolated to the full forward."""
    if rank != 0:
        return
    F_, H_, W_ = wl["grid"]
    T = F_ * H_ * W_
    forwards = 2 if wl["cfg"] != 1.0 else 1
    s = CpuSample(T, wl["Tc"])
    for _ in range(max(args.warmup, 1)):
        s.block()
    times = [s.block() for _ in range(args.steps)]
    ms = (48 * statistics.mean(times) + s.t_prepost) * forwards * 1e3
    value = T / (ms / 1e3)
    cfg = workload_config(wl, 48, 2 if (forwards == 2 and args.gpus == 1) else 1, forwards, "host cpu", False)
    cfg["parallelism_ours"] = "see the other arm"
    line = {"impl": "reference", "metric": "video_tokens_per_s", "value": value, "unit": "tokens/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "steps_per_s": 1e3 / ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "extrapolated": True,
            "config": cfg,
            "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": s.cores, "kind": "port", "sample": s.describe(times), "extrapolated": True},
            "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), file=result_out, flush=True)


# --------------------------------------------------------------------------------------------------
def _json_only_stdout():
    """Stdout carries exactly ONE JSON line: everything else any library prints there (NCCL's version banner under
    torchrun, for one) is sent to stderr.  Returns the writer for the result line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def parity_check(M, par, audio: bool, use_cfg: bool, dev, world: int, rank: int) -> dict:
    """Checker, before any timed region: a 2-layer small model (production head widths) through THIS run's parallel
    layout against (a) the same model un-sharded on this GPU — must be bit-identical — and (b) the fp32 oracle on the
    host (rank 0): BASELINE bar rel-L2 <= 1e-2, cosine >= 0.999.  At N = 1 only (b)."""
    import torch
    import torch.distributed as dist

    sys.path.insert(0, str(ROOT / "oracle"))
    import ltx_oracle as O

    from mlx_video_b200 import sampler

    mt_o = O.LTXModelType.AudioVideo if audio else O.LTXModelType.VideoOnly
    heads = 8  # divisible by every Ulysses degree of one box
    cfg = O.small_config(mt_o, num_layers=2, heads=heads, audio_heads=heads)
    tensors = O.init_params(cfg, seed=0)
    tensors = {k: (v.to(torch.bfloat16).float() if k.endswith(".weight") and not k.endswith("_norm.weight") else v) for k, v in tensors.items()}
    d = {k: getattr(cfg, k) for k in cfg.__dataclass_fields__}
    d["model_type"], d["rope_type"] = d["model_type"].value, d["rope_type"].value
    pcfg = M.LTXModelConfig.from_dict(d)
    g = torch.Generator().manual_seed(4321)
    F_, H_, W_, Tc, Ta = 4, 8, 8, 40, 21
    T = F_ * H_ * W_
    lat, ctx_p, ctx_n = torch.randn(1, T, 128, generator=g), torch.randn(1, Tc, cfg.caption_channels, generator=g), torch.randn(1, Tc, cfg.caption_channels, generator=g)
    pos = torch.from_numpy(sampler.create_position_grid(1, F_, H_, W_))
    ts = torch.full((1, T), 0.725)
    a_lat, a_ctx = torch.randn(1, Ta, 128, generator=g), torch.randn(1, Tc, cfg.audio_caption_channels, generator=g)
    a_pos, a_ts = torch.from_numpy(sampler.create_audio_position_grid(1, Ta)), torch.full((1, Ta), 0.725)

    def mods(ctx):
        v = M.Modality(lat.to(dev), ts.to(dev), pos.to(dev), ctx.to(dev))
        a = M.Modality(a_lat.to(dev), a_ts.to(dev), a_pos.to(dev), a_ctx.to(dev)) if audio else None
        return v, a

    single = M.LTXModel(pcfg, device=dev)
    single.load_weights(tensors)
    refs = [single(*mods(c)) for c in ((ctx_p, ctx_n) if use_cfg else (ctx_p,))]
    out = {"model": f"2 layers, {heads} heads x 128 (audio {heads} x 64), T={T}, Tc={Tc}" + (f", Ta={Ta}" if audio else "")}
    if par is not None:
        sharded = M.LTXModel(pcfg, device=dev)
        sharded.load_weights(tensors)
        par.attach(sharded)
        def rel(a, b):
            return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))

        if par.cfg is not None:
            mine, _ = sharded(*mods(ctx_p if par.cfg.is_cond else ctx_n))
            got = list(par.cfg.exchange(mine))
            pairs = [(g_, r_[0]) for g_, r_ in zip(got, refs)]
        else:
            gv, ga = sharded(*mods(ctx_p))
            pairs = [(gv, refs[0][0])] + ([(ga, refs[0][1])] if audio else [])
        same = all(torch.equal(g_, r_) for g_, r_ in pairs)
        worst = max(rel(g_, r_) for g_, r_ in pairs)
        stats = torch.tensor([1.0 if same else 0.0, -worst], device=dev, dtype=torch.float64)
        dist.all_reduce(stats, op=dist.ReduceOp.MIN)
        # bit identity holds when the shards happen to take the same tile / key-split schedule as the whole problem; what is
        # required is agreement to summation order: rel-L2 <= 5e-3 against the un-sharded model on the same GPU
        out["bit_identical"] = bool(stats[0].item() > 0.5)
        out["rel_l2_vs_single_gpu"] = float(-stats[1].item())
        out["layout"] = par.describe()
    if rank == 0:
        torch.set_num_threads(os.cpu_count() or 1)
        with torch.no_grad():
            wv, wa = O.OracleLTXModel(cfg, tensors)(O.Modality(lat, ts, pos, ctx_p), O.Modality(a_lat, a_ts, a_pos, a_ctx) if audio else None)
        gv = refs[0][0].cpu().double()
        out["rel_l2_vs_oracle"] = float((gv - wv.double()).norm() / wv.double().norm())
        out["cosine_vs_oracle"] = float(torch.nn.functional.cosine_similarity(gv.flatten(), wv.double().flatten(), dim=0))
        if audio:
            ga = refs[0][1].cpu().double()
            out["audio_rel_l2_vs_oracle"] = float((ga - wa.double()).norm() / wa.double().norm())
        out["ok"] = bool(out["rel_l2_vs_oracle"] <= 1e-2 and out["cosine_vs_oracle"] >= 0.999 and out.get("rel_l2_vs_single_gpu", 0.0) <= 5e-3)
    del single
    torch.cuda.empty_cache()
    return out


def run_workload(name: str, wl: dict, args, ctx: dict, models: dict, steps: int, warmup: int, headline: bool) -> dict:
    """Time one workload: device-resident, end-to-end through the public API, per-kernel leg.  Returns rank 0's record
    (other ranks: {})."""
    import torch
    import torch.distributed as dist

    import mlx_video_b200 as M
    from mlx_video_b200 import ops, sampler

    world, rank, dev, pk = ctx["world"], ctx["rank"], ctx["dev"], ctx["peaks"]
    F_, H_, W_ = wl["grid"]
    T, Tc, cfg_scale = F_ * H_ * W_, wl["Tc"], wl["cfg"]
    use_cfg = cfg_scale != 1.0
    Ta = wl.get("Ta", 0)

    # ---------------- parallel layout
    par = None
    parallelism = "single"
    if world > 1:
        from mlx_video_b200 import parallel

        par = parallel.make_layout(world, rank, use_cfg, fused=False if args.nccl_exchange else None)
        parallelism = par.describe()
    parity = parity_check(M, par, bool(Ta), use_cfg, dev, world, rank) if not args.no_parity else None

    # ---------------- model + resident inputs (synthetic, seeded; random-init weights of the real architecture)
    # graph replay: single-GPU, and multi-GPU when the exchange is NVLink-fused (capturing NCCL collectives hung)
    use_graph = (not args.no_graph) and (world == 1 or not args.nccl_exchange)
    cache = not args.no_cache_context
    mkey = "av" if Ta else "video"
    if mkey not in models:
        for k in list(models):  # one 19B-parameter model resident at a time
            del models[k]
        torch.cuda.empty_cache()
        models[mkey] = M.LTXModel(M.production_config(M.LTXModelType.AudioVideo if Ta else M.LTXModelType.VideoOnly, num_layers=args.layers),
                                  device=dev, cuda_graphs=use_graph, cache_context=cache).init_random(seed=0)
    model = models[mkey]
    model.clear_caches()
    model.seq_parallel = None
    if par is not None:
        par.attach(model)
    g = torch.Generator().manual_seed(1234)
    lat_h = torch.randn(1, T, 128, generator=g).pin_memory()
    ctx_pos_h = torch.randn(1, Tc, 3840, generator=g).to(torch.bfloat16).pin_memory()
    ctx_neg_h = torch.randn(1, Tc, 3840, generator=g).to(torch.bfloat16).pin_memory()
    pos_h = torch.from_numpy(sampler.create_position_grid(1, F_, H_, W_)).pin_memory()
    ones_h = torch.ones(1, T).pin_memory()
    sig = sampler.STAGE_1_SIGMAS if not use_cfg else [float(s) for s in sampler.ltx2_scheduler(40, T)]
    if name == "long":
        sig = [1.0] + list(sampler.STAGE_2_SIGMAS)  # stage 2 has 3 steps; one more leading sigma so short runs see 3 distinct ones
    n_sig = len(sig) - 2  # never take the final step to sigma=0 inside the loop (x would collapse to x0 and stay)

    x0 = lat_h.to(dev)
    x = x0.clone()
    ctx_pos, ctx_neg, pos = ctx_pos_h.to(dev), ctx_neg_h.to(dev), pos_h.to(dev)
    ones = ones_h.to(dev)
    rope = sampler._video_rope(model, pos)
    b = 2 if (use_cfg and world == 1) else 1  # cfg_batch on one GPU
    if b == 2:
        ctx_cat = torch.cat([ctx_pos, ctx_neg], 0)
        pos_cat = torch.cat([pos, pos], 0)
    ts_buf = torch.empty(b, T, device=dev)
    if Ta:  # audio stream: latents (1, Ta, 128), its own text context, positions in seconds
        a0 = torch.randn(1, Ta, 128, generator=g).to(dev)
        xa = a0.clone()
        a_ctx = torch.randn(1, Tc, 3840, generator=g).to(torch.bfloat16).to(dev)
        a_pos = torch.from_numpy(sampler.create_audio_position_grid(1, Ta)).to(dev)
        a_rope = sampler._audio_rope(model, a_pos)
        a_ones = torch.ones(1, Ta, device=dev)
    state = {}

    def forward(xin, sigma):
        """One denoise step's model work -> (v_pos, v_neg)."""
        if b == 2:
            ts_buf.copy_(ones.expand(2, T) * sigma)
            vv, _ = model(video=M.Modality(torch.cat([xin, xin], 0), ts_buf, pos_cat, ctx_cat, True, None, rope), audio=None)
            return vv[:1], vv[1:]
        ts_buf.copy_(ones * sigma)
        if par is not None and par.cfg is not None:
            c = ctx_pos if par.cfg.is_cond else ctx_neg
            mine, _ = model(video=M.Modality(xin, ts_buf, pos, c, True, None, rope), audio=None)
            return par.cfg.exchange(mine)
        am = M.Modality(xa, a_ones * sigma, a_pos, a_ctx, True, None, a_rope) if Ta else None
        v, va = model(video=M.Modality(xin, ts_buf, pos, ctx_pos, True, None, rope), audio=am)
        if Ta:
            state["audio_velocity"] = va
        return v, None

    def step(i):
        k = i % n_sig
        if k == 0:  # a new denoise loop: fresh latents and a fresh prompt (its text K/V are projected once, on this step)
            x.copy_(x0)
            model.invalidate_context()
        v_pos, v_neg = forward(x, sig[k])
        sampler._advance(x, v_pos.contiguous(), sig[k], sig[k + 1], v_neg=None if v_neg is None else v_neg.contiguous(), cfg_scale=cfg_scale)
        if Ta:
            if k == 0:
                xa.copy_(a0)
            sampler._advance(xa, state["audio_velocity"].contiguous(), sig[k], sig[k + 1])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing
    for i in range(warmup):
        step(i)
    barrier()
    model.check_timestep_groups()
    clocks = ClockSampler(ctx["local_rank"])
    clocks.start()
    launches0 = ops.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    gpu_launches = ops.launches - launches0  # host-side launches only: kernels replayed from a CUDA graph are recounted from an eager step below
    clk = clocks.stop()
    assert torch.isfinite(x).all(), "latents went non-finite"

    # ---------------- end to end: pinned host buffers in, velocity read back, every step.  Latents, per-token timesteps
    # and positions arrive from the host EVERY step; the text context is an input of the denoise LOOP (the reference's
    # loops take the embeddings once, generate.py:564-575) and is uploaded at each loop start (every n_sig steps).
    out_h = torch.empty(b, T, 128).pin_memory()
    ctx_bytes = ctx_pos_h.numel() * 2 * (2 if use_cfg else 1)
    h2d = lat_h.numel() * 4 + ones_h.numel() * 4 + pos_h.numel() * 4 + ctx_bytes / n_sig
    d2h = out_h.numel() * 4
    e2e_ctx = {"p": torch.empty_like(ctx_pos), "n": torch.empty_like(ctx_neg) if use_cfg else None, "cat": None}

    def e2e_step(i):
        sigma = sig[i % n_sig]
        if i % n_sig == 0:  # a new denoise loop: its prompt embeddings come from the host
            e2e_ctx["p"].copy_(ctx_pos_h, non_blocking=True)
            if use_cfg:
                e2e_ctx["n"].copy_(ctx_neg_h, non_blocking=True)
            if b == 2:
                e2e_ctx["cat"] = torch.cat([e2e_ctx["p"], e2e_ctx["n"]], 0)
        xin = lat_h.to(dev, non_blocking=True)
        ts_dev = ones_h.to(dev, non_blocking=True) * sigma  # timesteps = sigma * mask, as the reference's loop builds them
        p = pos_h.to(dev, non_blocking=True)
        if b == 2:
            m = M.Modality(torch.cat([xin, xin], 0), torch.cat([ts_dev, ts_dev], 0), torch.cat([p, p], 0), e2e_ctx["cat"], True, None, rope)
            v, _ = model(video=m, audio=None)
        elif par is not None and par.cfg is not None:
            m = M.Modality(xin, ts_dev, p, e2e_ctx["p"] if par.cfg.is_cond else e2e_ctx["n"], True, None, rope)
            v, _ = model(video=m, audio=None)
        else:
            m = M.Modality(xin, ts_dev, p, e2e_ctx["p"], True, None, rope)
            am = M.Modality(a0, a_ones * sigma, a_pos, a_ctx, True, None, a_rope) if Ta else None
            v, _ = model(video=m, audio=am)
        out_h.copy_(v, non_blocking=True)
        torch.cuda.synchronize()

    for i in range(2):
        e2e_step(i)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    f0.record()
    for i in range(steps):
        e2e_step(i)
    f1.record()
    barrier()
    e2e_wall_ms = (time.perf_counter() - t_wall0) * 1e3
    e2e_ms = max(f0.elapsed_time(f1), 0.0)

    # ---------------- per-kernel timing (CUDA events around every launch, same stream) for the roofline leg
    graphs, model._graphs = model._graphs, None  # eager for this leg: the brackets sit between the launches
    fill_prof = ops.profile(True)
    torch.cuda._sleep(int(3e7))  # give the host a head start so launch latency is not inside the brackets
    step(0)  # the first step of a loop: a steady step's work + the once-per-loop context projections
    torch.cuda.synchronize()
    ops.profile(False)
    launches_step0 = ops.launches
    prof = ops.profile(True)
    torch.cuda._sleep(int(3e7))
    step(1)  # a steady step of the loop (text K/V of this loop's prompt already projected when the cache is on)
    torch.cuda.synchronize()
    gemm_shapes = list(ops._profile_shapes)  # (M, N, K) per GEMM launch of the steady step
    ops.profile(False)
    launches_per_step = ops.launches - launches_step0
    model._graphs = graphs
    table = {}
    for kname, work, a, bb in prof:
        t = table.setdefault(kname, dict(launches=0, ms=0.0, work=0.0))
        t["launches"] += 1
        t["ms"] += a.elapsed_time(bb)
        t["work"] += work
    prof_total = sum(t["ms"] for t in table.values())
    shapes = {}  # GEMM launches grouped by their FLOP count (= by problem shape), for --kernel-table
    for kname, work, a, bb in prof:
        if kname == "ltxb_gemm_bf16":
            t = shapes.setdefault(work, [0, 0.0])
            t[0] += 1
            t[1] += a.elapsed_time(bb)
    contract = ("ltxb_gemm_bf16", "ltxb_attention_fwd", "ltxb_attention_fwd_peers")
    steady_flops = sum(t["work"] for kname, t in table.items() if kname in contract)
    first_flops = sum(work for kname, work, _, _ in fill_prof if kname in contract)
    executed_flops = (first_flops + (n_sig - 1) * steady_flops) / n_sig  # mean over one denoise loop

    # ---------------- reduce over ranks (max time; executed work summed)
    times = torch.tensor([ms_total, e2e_ms, e2e_wall_ms], device=dev, dtype=torch.float64)
    work_t = torch.tensor([executed_flops], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(work_t, op=dist.ReduceOp.SUM)
    ms_total, e2e_ms, e2e_wall_ms = [float(v) for v in times.tolist()]
    executed_flops = float(work_t.item())
    if par is not None:
        model.seq_parallel = None
    if rank != 0:
        return {}

    ms_step = ms_total / steps
    forwards = 2 if use_cfg else 1
    flops_step = model.forward_flops(T, Tc, Ta) * forwards
    tflops_alg = flops_step / (ms_step * 1e-3) / 1e12
    tflops_exec = executed_flops / (ms_step * 1e-3) / 1e12
    gemm = table.get("ltxb_gemm_bf16", dict(launches=1, ms=1e-9, work=0.0))
    gemm_tflops = gemm["work"] / (gemm["ms"] * 1e-3) / 1e12
    traffic = None
    tpath = ROOT / "profiles" / "roofline_traffic.json"
    if tpath.exists():
        traffic = json.loads(tpath.read_text()).get("ltxb_gemm_bf16")
    roofline = {"kernel": "gemm_bf16_kernel (tcgen05/TMEM, TMA-fed)", "bound": "tensor", "achieved": gemm_tflops, "peak": pk["tflops"],
                "unit": "TFLOP/s", "frac": gemm_tflops / pk["tflops"], "traffic": traffic, "peak_source": pk["source"] + " sustained",
                "launches_per_step": gemm["launches"], "share_of_step": gemm["ms"] / max(prof_total, 1e-9)}
    # Few rows per rank (a sequence-parallel shard of <= 512 tokens: every launch goes to gemm_small_m_kernel): the weight
    # stream from HBM bounds the GEMMs, not the tensor pipe — algorithmic bytes = every launch's N x K bf16 weights.
    few = [s for s in gemm_shapes if s[0] <= 512]
    if gemm_shapes and sum(2.0 * n * k for _, n, k in few) >= 0.9 * sum(2.0 * n * k for _, n, k in gemm_shapes):
        w_bytes = sum(2.0 * n * k for _, n, k in gemm_shapes)
        gbs = w_bytes / (gemm["ms"] * 1e-3) / 1e9
        roofline = {"kernel": "gemm_small_m_kernel (weight rows on the TMEM lanes, tcgen05, TMA-fed)", "bound": "hbm", "achieved": gbs,
                    "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"], "traffic": w_bytes / max(gemm["launches"], 1),
                    "traffic_source": "algorithmic (ncu of the 160 x 16384 x 4096 launch: 135.7 MB read for 134.2 MB of weights, profiles/r2/gemm/few_row_kernel_ncu_160x16384x4096.json)",
                    "peak_source": pk["source"] + " HBM copy", "launches_per_step": gemm["launches"],
                    "share_of_step": gemm["ms"] / max(prof_total, 1e-9), "tensor_tflops": gemm_tflops}
    rec = {
        "metric": "video_tokens_per_s", "value": T * steps / (ms_total * 1e-3), "unit": "tokens/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": ms_step, "steps_per_s": 1e3 / ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(wl, args.layers, b, forwards, parallelism, cache),
        # executed work = what the kernels of one step actually contracted (the text K/V projection runs once per loop when
        # cached, AdaLN on distinct timesteps only); algorithmic = the reference's un-deduplicated count
        "model_tflops": tflops_exec, "model_tflops_algorithmic": tflops_alg,
        "pct_of_bf16_peak": {"sustained": 100 * tflops_exec / (pk["tflops"] * world), "burst": 100 * tflops_exec / (pk["burst"] * world),
                             "on": "executed work", "peaks": pk["source"]},
        "algorithmic_tflop_per_step": flops_step / 1e12, "executed_tflop_per_step": executed_flops / 1e12,
        "clocks": clk,
        "e2e": {"value": T * steps / (e2e_ms * 1e-3), "unit": "tokens/s", "ms_per_step": e2e_ms / steps,
                "wall_ms_per_step": e2e_wall_ms / steps, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": d2h,
                "context_upload": f"once per {n_sig}-step denoise loop ({ctx_bytes} bytes), amortised into h2d_bytes_per_step"},
        "gpu_launches": gpu_launches if not use_graph else launches_per_step * steps,
        "launch_mode": "eager" if not use_graph else f"cuda graph replay ({launches_per_step} kernels per step captured)",
        "context_cache": cache,
        "roofline": roofline,
        "kernels": {k: {"launches": v["launches"], "ms": round(v["ms"], 4), "share": round(v["ms"] / max(prof_total, 1e-9), 4)} for k, v in sorted(table.items(), key=lambda kv: -kv[1]["ms"])},
    }
    if world > 1:
        rec["parallel_parity"] = parity
    else:
        rec["parity"] = parity
    for aname in ("ltxb_attention_fwd", "ltxb_attention_fwd_peers"):
        att = table.get(aname)
        if att and att["ms"] > 0:
            rec.setdefault("attention_tflops", {})[aname] = att["work"] / (att["ms"] * 1e-3) / 1e12
    if headline and not args.no_cpu_baseline:
        s = cpu_block_sample(T, Tc)  # video stream of the workload (the audio stream adds 12 % of the FLOPs on top)
        cpu_value = T / (s["t_forward"] * forwards)
        rec["cpu_baseline"] = {"value": cpu_value, "unit": "tokens/s", "cores": s["cores"], "kind": "port", "sample": s["sample"], "extrapolated": True}
    if args.kernel_table:
        print(f"--- {name}: {ms_step:.3f} ms/step ({parallelism})", file=sys.stderr)
        for k, v in rec["kernels"].items():
            print(f"{k:34s} {v['launches']:5d} launches {v['ms']:9.3f} ms {100 * v['share']:5.1f}%", file=sys.stderr)
        for work, (n, ms) in sorted(shapes.items(), key=lambda kv: -kv[1][1]):
            print(f"  gemm {work / 1e9:9.1f} GFLOP x {n:3d}: {ms / n * 1e3:8.1f} us each, {work * n / ms / 1e9:7.0f} TFLOP/s, {ms:7.3f} ms", file=sys.stderr)
    if not headline:  # compact record for the "workloads" object
        keep = ("value", "unit", "ms_per_step", "steps_per_s", "steps", "warmup", "model_tflops", "model_tflops_algorithmic", "pct_of_bf16_peak",
                "algorithmic_tflop_per_step", "executed_tflop_per_step", "e2e", "launch_mode", "roofline", "attention_tflops", "parallel_parity", "parity")
        rec = {"config": rec["config"], **{k: rec[k] for k in keep if k in rec}}
    return rec


def main() -> int:
    result_out = _json_only_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="distilled", choices=sorted(WORKLOADS), help="the headline workload of the JSON line")
    ap.add_argument("--workloads", default="dev,av,long", help="other BASELINE configs measured briefly into \"workloads\" (comma list, or none)")
    ap.add_argument("--sub-steps", type=int, default=3, help="timed steps of each entry of --workloads (3 warm-up steps each)")
    ap.add_argument("--layers", type=int, default=48, help="debug only: fewer blocks (the JSON line says so)")
    ap.add_argument("--text-tokens", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the small-model parity check that precedes the timed region")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-cache-context", action="store_true", help="recompute the caption projection and text K/V every step like the reference (default: once per denoise loop, row N1)")
    ap.add_argument("--cache-context", action="store_true", help="(default; kept for older scripts)")
    ap.add_argument("--nccl-exchange", action="store_true", help="sequence parallelism through NCCL all_to_all instead of the NVLink-fused kernels")
    ap.add_argument("--kernel-table", action="store_true", help="print the per-kernel time table to stderr")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = dict(WORKLOADS[args.workload])
    if args.text_tokens:
        wl["Tc"] = args.text_tokens
    if args.impl == "reference":
        run_reference(args, wl, rank, result_out)
        return 0
    if world != args.gpus:
        if args.gpus != 1 and world == 1:
            print(f"bench.py: --gpus {args.gpus} needs a torchrun launch with {args.gpus} ranks", file=sys.stderr)
            return 2

    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = dict(world=world, rank=rank, local_rank=local_rank, dev=dev, peaks=peaks())
    models: dict = {}
    line = run_workload(args.workload, wl, args, ctx, models, args.steps, args.warmup, headline=True)
    subs = [w for w in args.workloads.split(",") if w and w != "none" and w != args.workload]
    others = {}
    for name in subs:
        if name not in WORKLOADS:
            print(f"bench.py: unknown workload {name!r} in --workloads", file=sys.stderr)
            return 2
        t0 = time.perf_counter()
        try:
            others[name] = run_workload(name, dict(WORKLOADS[name]), args, ctx, models, args.sub_steps, 3, headline=False)
        except Exception as e:  # a failing side workload must not take the headline line with it
            others[name] = {"error": f"{type(e).__name__}: {e}"[:400]}
            if world > 1:
                raise
        if rank == 0:
            others[name]["bench_wall_s"] = round(time.perf_counter() - t0, 1)
    if rank == 0:
        if others:
            line["workloads"] = others
        print(json.dumps(line), file=result_out, flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
