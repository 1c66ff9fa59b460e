#!/usr/bin/env python
"""LTX-2 DiT denoise-step throughput on B200 (BASELINE.json metric: video tokens/s and steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload distilled|dev]

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
One JSON line on stdout (rank 0).  `value` has inputs resident in HBM; `e2e` goes through the public
API with pinned HOST buffers copied in and the velocity read back every step.
The only place this file touches oracle/ is the CPU baseline (`cpu_baseline`, `--impl reference`).
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
                f"{statistics.median(times) * 1e3:.0f} ms, + pre/post {self.t_prepost * 1e3:.0f} ms; forward = 48 x block + pre/post")


def cpu_block_sample(T: int, Tc: int, budget_s: float = 15.0) -> dict:
    s = CpuSample(T, Tc)
    s.block()  # warm-up
    times, t_start = [], time.perf_counter()
    while len(times) < 2 or (time.perf_counter() - t_start < budget_s and len(times) < 8):
        times.append(s.block())
    return dict(t_forward=48 * statistics.median(times) + s.t_prepost, cores=s.cores, sample=s.describe(times))


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
    line = {"impl": "reference", "metric": "video_tokens_per_s", "value": value, "unit": "tokens/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "steps_per_s": 1e3 / ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "video_tokens": T, "text_tokens": wl["Tc"], "forwards_per_step": forwards},
            "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": s.cores, "kind": "port", "sample": s.describe(times)},
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


def main() -> int:
    result_out = _json_only_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="distilled", choices=sorted(WORKLOADS))
    ap.add_argument("--layers", type=int, default=48, help="debug only: fewer blocks (the JSON line says so)")
    ap.add_argument("--text-tokens", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--cache-context", action="store_true", help="headline run WITH the cross-step text K/V cache (row N1); default recomputes them like the reference")
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

    import mlx_video_b200 as M
    from mlx_video_b200 import ops, sampler

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pk = peaks()
    F_, H_, W_ = wl["grid"]
    T, Tc, cfg_scale = F_ * H_ * W_, wl["Tc"], wl["cfg"]
    use_cfg = cfg_scale != 1.0

    # ---------------- parallel layout
    par = None
    parallelism = "single"
    if world > 1:
        from mlx_video_b200 import parallel

        par = parallel.make_layout(world, rank, use_cfg, fused=False if args.nccl_exchange else None)
        parallelism = par.describe()

    # ---------------- model + resident inputs (synthetic, seeded; random-init weights of the real architecture)
    # graph replay is single-GPU only for now: capturing the NCCL exchanges of the sequence-parallel path hung on the
    # first attempt (profiles/README.md), so multi-rank runs launch eagerly
    use_graph = (not args.no_graph) and (world == 1 or not args.nccl_exchange)
    Ta = wl.get("Ta", 0)
    model = M.LTXModel(M.production_config(M.LTXModelType.AudioVideo if Ta else M.LTXModelType.VideoOnly, num_layers=args.layers), device=dev,
                       cuda_graphs=use_graph, cache_context=args.cache_context).init_random(seed=0)
    if par is not None:
        par.attach(model)
    g = torch.Generator().manual_seed(1234)
    lat_h = torch.randn(1, T, 128, generator=g).pin_memory()
    ctx_pos_h = torch.randn(1, Tc, 3840, generator=g).to(torch.bfloat16).pin_memory()
    ctx_neg_h = torch.randn(1, Tc, 3840, generator=g).to(torch.bfloat16).pin_memory()
    pos_h = torch.from_numpy(sampler.create_position_grid(1, F_, H_, W_)).pin_memory()
    ones_h = torch.ones(1, T).pin_memory()
    sig = sampler.STAGE_1_SIGMAS if not use_cfg else [float(s) for s in sampler.ltx2_scheduler(40, T)]
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

    def forward(xin, sigma):
        """One denoise step's model work -> (v_pos, v_neg)."""
        if b == 2:
            ts_buf.copy_(ones.expand(2, T) * sigma)
            vv, _ = model(video=M.Modality(torch.cat([xin, xin], 0), ts_buf, pos_cat, ctx_cat, True, None, rope), audio=None)
            return vv[:1], vv[1:]
        ts_buf.copy_(ones * sigma)
        if par is not None and par.cfg is not None:
            ctx = ctx_pos if par.cfg.is_cond else ctx_neg
            mine, _ = model(video=M.Modality(xin, ts_buf, pos, ctx, True, None, rope), audio=None)
            return par.cfg.exchange(mine)
        am = M.Modality(xa, a_ones * sigma, a_pos, a_ctx, True, None, a_rope) if Ta else None
        v, va = model(video=M.Modality(xin, ts_buf, pos, ctx_pos, True, None, rope), audio=am)
        if Ta:
            forward.audio_velocity = va
        return v, None

    def step(i):
        k = i % n_sig
        if k == 0:
            x.copy_(x0)
        v_pos, v_neg = forward(x, sig[k])
        sampler._advance(x, v_pos.contiguous(), sig[k], sig[k + 1], v_neg=None if v_neg is None else v_neg.contiguous(), cfg_scale=cfg_scale)
        if Ta:
            if k == 0:
                xa.copy_(a0)
            sampler._advance(xa, forward.audio_velocity.contiguous(), sig[k], sig[k + 1])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing
    for i in range(args.warmup):
        step(i)
    barrier()
    model.check_timestep_groups()
    clocks = ClockSampler(local_rank)
    clocks.start()
    launches0 = ops.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    gpu_launches = ops.launches - launches0  # host-side launches only: kernels replayed from a CUDA graph are recounted from an eager step below
    clk = clocks.stop()
    assert torch.isfinite(x).all(), "latents went non-finite"

    # ---------------- end to end: pinned host buffers in, velocity read back, every step
    out_h = torch.empty(b, T, 128).pin_memory()
    h2d = lat_h.numel() * 4 + ones_h.numel() * 4 + pos_h.numel() * 4 + ctx_pos_h.numel() * 2 * (2 if use_cfg else 1)
    d2h = out_h.numel() * 4

    def e2e_step(i):
        sigma = sig[i % n_sig]
        xin = lat_h.to(dev, non_blocking=True)
        ts_dev = ones_h.to(dev, non_blocking=True) * sigma  # timesteps = sigma * mask, as the reference's loop builds them
        p = pos_h.to(dev, non_blocking=True)
        cp = ctx_pos_h.to(dev, non_blocking=True)
        cn = ctx_neg_h.to(dev, non_blocking=True) if use_cfg else None
        if b == 2:
            m = M.Modality(torch.cat([xin, xin], 0), torch.cat([ts_dev, ts_dev], 0), torch.cat([p, p], 0), torch.cat([cp, cn], 0), True, None, rope)
            v, _ = model(video=m, audio=None)
            out_h.copy_(v, non_blocking=True)
        elif par is not None and par.cfg is not None:
            m = M.Modality(xin, ts_dev, p, cp if par.cfg.is_cond else cn, True, None, rope)
            v, _ = model(video=m, audio=None)
            out_h.copy_(v, non_blocking=True)
        else:
            m = M.Modality(xin, ts_dev, p, cp, True, None, rope)
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
    for i in range(args.steps):
        e2e_step(i)
    f1.record()
    barrier()
    e2e_wall_ms = (time.perf_counter() - t_wall0) * 1e3
    e2e_ms = max(f0.elapsed_time(f1), 0.0)

    # ---------------- per-kernel timing (CUDA events around every launch, same stream) for the roofline leg
    graphs, model._graphs = model._graphs, None  # eager for this leg: the brackets sit between the launches
    launches_step0 = ops.launches
    prof = ops.profile(True)
    torch.cuda._sleep(int(3e7))  # give the host a head start so launch latency is not inside the brackets
    step(1)
    torch.cuda.synchronize()
    ops.profile(False)
    launches_per_step = ops.launches - launches_step0
    model._graphs = graphs
    table = {}
    for name, work, a, bb in prof:
        t = table.setdefault(name, dict(launches=0, ms=0.0, work=0.0))
        t["launches"] += 1
        t["ms"] += a.elapsed_time(bb)
        t["work"] += work
    prof_total = sum(t["ms"] for t in table.values())
    shapes = {}  # GEMM launches grouped by their FLOP count (= by problem shape), for --kernel-table
    for name, work, a, bb in prof:
        if name == "ltxb_gemm_bf16":
            t = shapes.setdefault(work, [0, 0.0])
            t[0] += 1
            t[1] += a.elapsed_time(bb)

    # ---------------- reduce over ranks (max time)
    times = torch.tensor([ms_total, e2e_ms, e2e_wall_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, e2e_wall_ms = [float(v) for v in times.tolist()]

    if rank == 0:
        ms_step = ms_total / args.steps
        forwards = 2 if use_cfg else 1
        flops_step = model.forward_flops(T, Tc, Ta) * forwards
        tflops = flops_step / (ms_step * 1e-3) / 1e12
        gemm = table.get("ltxb_gemm_bf16", dict(launches=1, ms=1e-9, work=0.0))
        gemm_tflops = gemm["work"] / (gemm["ms"] * 1e-3) / 1e12
        traffic = None
        tpath = ROOT / "profiles" / "roofline_traffic.json"
        if tpath.exists():
            traffic = json.loads(tpath.read_text()).get("ltxb_gemm_bf16")
        line = {
            "metric": "video_tokens_per_s", "value": T * args.steps / (ms_total * 1e-3), "unit": "tokens/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "steps_per_s": 1e3 / ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": wl["desc"], "video_tokens": T, "audio_tokens": Ta, "text_tokens": Tc, "layers": args.layers, "batch": b,
                       "forwards_per_step": forwards, "parallelism": parallelism,
                       "l2": "48 blocks x 537 MB of bf16 weights stream through the 126 MB L2 every step (inputs larger than L2)",
                       "weights": "random-init, seeded, bf16; fp32 residual stream",
                       "text_kv": "cached across steps (row N1)" if args.cache_context else "recomputed every step, as the reference does"},
            "model_tflops": tflops, "pct_of_bf16_peak": {"sustained": 100 * tflops / (pk["tflops"] * world), "burst": 100 * tflops / (pk["burst"] * world), "peaks": pk["source"]},
            "algorithmic_tflop_per_step": flops_step / 1e12,
            "clocks": clk,
            "e2e": {"value": T * args.steps / (e2e_ms * 1e-3), "unit": "tokens/s", "ms_per_step": e2e_ms / args.steps,
                    "wall_ms_per_step": e2e_wall_ms / args.steps, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": gpu_launches if not use_graph else launches_per_step * args.steps,
            "launch_mode": "eager" if not use_graph else f"cuda graph replay ({launches_per_step} kernels per step captured)",
            "context_cache": bool(args.cache_context),
            "roofline": {"kernel": "gemm_bf16_kernel (tcgen05/TMEM, TMA-fed)", "bound": "tensor", "achieved": gemm_tflops, "peak": pk["tflops"],
                         "unit": "TFLOP/s", "frac": gemm_tflops / pk["tflops"], "traffic": traffic, "peak_source": pk["source"] + " sustained",
                         "launches_per_step": gemm["launches"], "share_of_step": gemm["ms"] / max(prof_total, 1e-9)},
            "kernels": {k: {"launches": v["launches"], "ms": round(v["ms"], 4), "share": round(v["ms"] / max(prof_total, 1e-9), 4)} for k, v in sorted(table.items(), key=lambda kv: -kv[1]["ms"])},
        }
        att = table.get("ltxb_attention_fwd")
        if att:
            line["attention_tflops"] = att["work"] / (att["ms"] * 1e-3) / 1e12
        if world == 1 and not args.no_cpu_baseline and not Ta:
            s = cpu_block_sample(T, Tc)
            cpu_value = T / (s["t_forward"] * forwards)
            line["cpu_baseline"] = {"value": cpu_value, "unit": "tokens/s", "cores": s["cores"], "kind": "port", "sample": s["sample"]}
        if args.kernel_table:
            for k, v in line["kernels"].items():
                print(f"{k:28s} {v['launches']:5d} launches {v['ms']:9.3f} ms {100 * v['share']:5.1f}%", file=sys.stderr)
            for work, (n, ms) in sorted(shapes.items(), key=lambda kv: -kv[1][1]):
                print(f"  gemm {work / 1e9:9.1f} GFLOP x {n:3d}: {ms / n * 1e3:8.1f} us each, {work * n / ms / 1e9:7.0f} TFLOP/s, {ms:7.3f} ms", file=sys.stderr)
        print(json.dumps(line), file=result_out, flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
