#!/bin/bash
# attention kernel tests + sweeps (+ variants, trace); Usage: bash scripts/gpu_r2b.sh <tag> [variants...]
tag=${1:-r2b}; shift; out=gpurun_out/$tag; mkdir -p $out
variants=${@:-"emu0 emu1 emu3 noil"}
echo "--- attention kernel tests, S64 (default)"
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k attention > $out/pytest_attn_s64.log 2>&1; rc=$?; echo "rc=$rc"; tail -5 $out/pytest_attn_s64.log
grep -E "rel_l2 [0-9]" $out/pytest_attn_s64.log | head
echo "--- sweeps"
timeout 200 python scripts/attn_sweep.py 2>&1 | tee $out/attn_sweep_s64.txt
for v in $variants; do
  echo $v; [ -f mlx-video_b200/csrc/libltxb_$v.so ] && LTXB_LIB=$PWD/mlx-video_b200/csrc/libltxb_$v.so ATTN_SWEEP_NO_SDPA=1 timeout 120 python scripts/attn_sweep.py 2>&1 | tee $out/attn_sweep_s64_$v.txt
done
for shape in "1 5184 5184 32 128" "1 1280 1280 32 128"; do
  LTXB_LIB=$PWD/mlx-video_b200/csrc/libltxb_trace.so timeout 60 python scripts/attn_trace.py $shape > "$out/attn_trace_$(echo $shape | tr ' ' '_').txt" 2>&1
done
