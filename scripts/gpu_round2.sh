#!/bin/bash
# One gpurun call that re-validates HEAD and refreshes the evidence under gpurun_out/<tag>/ (round 2):
#   pytest -m gpu, bench.py (headline + workloads, kernel table), attention / GEMM sweeps, VAE decode bench, then the ncu launch
#   list of one eager step and --set full captures of the top kernels (B200_PROFILING.md recipe).
# Usage: bash scripts/gpu_round2.sh <tag> [skip-ncu]
tag=${1:-r2}; out=gpurun_out/$tag; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > $out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest.log
timeout 900 python bench.py --kernel-table > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("$out/bench.json"))
print("bench", round(d["ms_per_step"], 3), "ms/step", round(d["value"]), "tok/s e2e", round(d["e2e"]["value"]), "exec TF", round(d["model_tflops"]),
      "gemm frac", round(d["roofline"]["frac"], 3), "attn TF", d.get("attention_tflops"), d["clocks"], d["launch_mode"])
print("cpu", d.get("cpu_baseline"))
for k, v in d.get("workloads", {}).items():
    print(k, {kk: (round(vv, 2) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("value", "ms_per_step", "model_tflops", "error")})
PY
grep -A24 "^--- distilled" $out/bench.err > $out/bench_kernel_table.txt; cat $out/bench_kernel_table.txt
timeout 300 python scripts/attn_sweep.py > $out/attn_sweep.txt 2>&1; cat $out/attn_sweep.txt
timeout 300 python scripts/gemm_sweep.py > $out/gemm_sweep.txt 2>&1; cat $out/gemm_sweep.txt
timeout 300 python scripts/gemm_small_m_bench.py 160x4096x4096 160x12288x4096 160x16384x4096 160x4096x16384 68x2048x2048 320x4096x4096 > $out/gemm_few_rows.txt 2>&1; cat $out/gemm_few_rows.txt
timeout 300 python bench.py --workload shard160 --workloads none --no-cpu-baseline --no-parity --steps 10 --warmup 3 --kernel-table > $out/bench_shard160.json 2> $out/bench_shard160.err; grep -A8 "^--- shard160" $out/bench_shard160.err
timeout 200 python scripts/vae_bench.py > $out/vae_bench.json 2>&1; cat $out/vae_bench.json
timeout 200 python scripts/upsampler_bench.py 2>&1 | tail -1 > $out/upsampler_bench.json; cat $out/upsampler_bench.json
[ "$2" = "skip-ncu" ] && exit 0
K='regex:^(gemm_bf16|attention|norm_modulate|qknorm_rope|gate_residual|timestep|rope_table|silu_bf16|cast_|euler_step)'
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph --workloads none --no-parity --no-cache-context"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base function -k "$K" -c 1520 --csv \
    --log-file $out/launches.csv $CMD > $out/ncu_launches.log 2>&1
echo "launch list rc=$?"
CMD2="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph --layers 4 --workloads none --no-parity --no-cache-context"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base function -k regex:^gemm_bf16 -s 34 -c 8 \
    -o $out/gemm $CMD2 > $out/ncu_gemm.log 2>&1
echo "gemm full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base function -k 'regex:^(attention|norm_modulate|qknorm_rope)' -s 20 -c 6 \
    -o $out/others $CMD2 > $out/ncu_others.log 2>&1
echo "others full rc=$?"
LTXB_BENCH_VARIANTS=small_m LTXB_BENCH_PACKED=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_small_m_kernel -s 10 -c 2 -o $out/gemm_few_rows python scripts/gemm_small_m_bench.py 160x16384x4096 > $out/ncu_gemm_few_rows.log 2>&1
echo "few-row gemm full rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_pair -s 2 -c 1 -o $out/attn_5184 python scripts/attn_one.py 1 5184 5184 32 128 5 > $out/ncu_attn_5184.log 2>&1
echo "attention 5184 full rc=$?"
ls -la $out
