#!/bin/bash
# ncu evidence for bench.py (B200_PROFILING.md recipe): launch list of one full step, then --set full on the
# top kernels.  Usage: bash scripts/ncu_bench.sh <tag>
tag=${1:-r1}; out=gpurun_out/ncu_$tag; mkdir -p $out
K='regex:^(gemm_bf16|attention|norm_modulate|qknorm_rope|gate_residual|timestep|rope_table|silu_bf16|cast_|euler_step)'
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > $out/plain.log 2>&1 || { echo "plain run failed"; tail -5 $out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base function -k "$K" -s 782 -c 782 --csv \
    --log-file $out/launches.csv $CMD > $out/ncu_launches.log 2>&1
echo "launch list rc=$?"
CMD2="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --layers 4"
$CMD2 > $out/plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base function -k regex:^gemm_bf16 -s 40 -c 7 \
    -o $out/gemm $CMD2 > $out/ncu_gemm.log 2>&1
echo "gemm full rc=$?"
ncu --set full --clock-control none --import-source on --kernel-name-base function -k 'regex:^(attention|norm_modulate|qknorm_rope)' -s 24 -c 5 \
    -o $out/others $CMD2 > $out/ncu_others.log 2>&1
echo "others full rc=$?"
ls -la $out
