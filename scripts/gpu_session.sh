#!/bin/bash
# Runs the bring-up groups, each in its own process under a timeout; logs to gpurun_out/<tag>/.
tag=${1:-s1}; shift
groups=${@:-"gemm_correct elementwise attn_correct gemm_perf attn_perf"}
out=gpurun_out/$tag; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > $out/smi.txt 2>&1
for g in $groups; do
  echo "=== $g ===" | tee -a $out/summary.txt
  timeout 300 python scripts/kernel_check.py $g > $out/$g.log 2> $out/$g.err
  rc=$?
  echo "$g rc=$rc" | tee -a $out/summary.txt
  grep -c '"ok": false' $out/$g.log | sed "s/^/$g failed checks: /" | tee -a $out/summary.txt
  tail -3 $out/$g.err | tee -a $out/summary.txt
done
grep -h '"ok": false' $out/*.log | head -40
grep -h '"perf"' $out/*.log | head -150
