#!/bin/bash
# ncu --set full of one small-M GEMM under two schedules (auto, data-parallel pair tiles). Usage: bash scripts/gpu_ncu_gemm_small.sh <tag>
out=gpurun_out/${1:-ncu_gemm_small}; mkdir -p $out
for sched in "-1 0" "1 256"; do
  set -- $sched
  name=pair$1_bn$2
  python scripts/gemm_one.py 160 4096 4096 $1 $2 > $out/plain_$name.log 2>&1 || { echo "plain failed $name"; tail -3 $out/plain_$name.log; continue; }
  cat $out/plain_$name.log
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 3 -c 1 -o $out/gemm_$name python scripts/gemm_one.py 160 4096 4096 $1 $2 > $out/ncu_$name.log 2>&1
  echo "ncu rc=$? ($name)"
done
ls -la $out
