#!/bin/bash
# ncu --set full of both attention kernels on one shape (plain runs first); Usage: bash scripts/gpu_ncu_attn2.sh <tag> "<B Tq Tk H dh>"
out=gpurun_out/${1:-ncu_attn2}; mkdir -p $out
args="${2:-1 5184 5184 32 128}"
for v in 1 0; do
  LTXB_ATTN_S64=$v python scripts/attn_one.py $args 5 > $out/plain_s64_$v.log 2>&1 || { echo "plain run failed (S64=$v)"; tail -5 $out/plain_s64_$v.log; continue; }
  cat $out/plain_s64_$v.log
  LTXB_ATTN_S64=$v timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_pair -s 2 -c 1 -o $out/attn_s64_$v python scripts/attn_one.py $args 5 > $out/ncu_s64_$v.log 2>&1
  echo "ncu rc=$? (S64=$v)"
done
ls -la $out
