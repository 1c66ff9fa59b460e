#!/bin/bash
# ncu launch list of one eager step in the few-row regime (160 video tokens on one GPU = one rank's GEMM rows of the 8-GPU run)
out=gpurun_out/r2ai; mkdir -p $out
CMD="python bench.py --workload shard160 --steps 1 --warmup 1 --no-cpu-baseline --no-graph --workloads none --no-parity --no-cache-context"
$CMD > $out/plain.json 2> $out/plain.err; echo "plain rc=$?"
K='regex:^(gemm_|attention|norm_modulate|qknorm_rope|gate_residual|timestep|rope_table|silu_bf16|cast_|euler_step)'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base function -k "$K" -c 1600 --csv --log-file $out/launches_shard160.csv $CMD > $out/ncu.log 2>&1
echo "launch list rc=$?"; wc -l $out/launches_shard160.csv
