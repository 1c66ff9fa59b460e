#!/bin/bash
# One gpurun call that re-validates HEAD and refreshes the evidence under gpurun_out/<tag>/:
#   pytest -m gpu, bench.py (default workload, 16 steps, kernel table), attention / GEMM sweeps, then the
#   ncu launch list of one eager step and --set full captures of the top kernels (B200_PROFILING.md recipe).
# Usage: bash scripts/gpu_round.sh <tag> [skip-ncu]
tag=${1:-r1b}; out=gpurun_out/$tag; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > $out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest.log
timeout 600 python bench.py --kernel-table > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
python - <<EOF
import json
d = json.load(open("$out/bench.json"))
print("bench", round(d["ms_per_step"], 3), "ms/step", round(d["value"]), "tok/s e2e", round(d["e2e"]["value"]), "model TF", round(d["model_tflops"]),
      "gemm frac", round(d["roofline"]["frac"], 3), "attn TF", round(d.get("attention_tflops", 0)), d["clocks"], d["launch_mode"])
print("cpu", d.get("cpu_baseline"))
EOF
tail -25 $out/bench.err
timeout 300 python scripts/attn_sweep.py > $out/attn_sweep.txt 2>&1; cat $out/attn_sweep.txt
[ "$2" = "skip-ncu" ] && exit 0
K='regex:^(gemm_bf16|attention|norm_modulate|qknorm_rope|gate_residual|timestep|rope_table|silu_bf16|cast_|euler_step)'
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base function -k "$K" -c 1520 --csv \
    --log-file $out/launches.csv $CMD > $out/ncu_launches.log 2>&1
echo "launch list rc=$?"
CMD2="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph --layers 4"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base function -k regex:^gemm_bf16 -s 34 -c 8 \
    -o $out/gemm $CMD2 > $out/ncu_gemm.log 2>&1
echo "gemm full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base function -k 'regex:^(attention|norm_modulate|qknorm_rope)' -s 20 -c 6 \
    -o $out/others $CMD2 > $out/ncu_others.log 2>&1
echo "others full rc=$?"
ls -la $out
