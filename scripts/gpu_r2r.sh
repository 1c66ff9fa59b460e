#!/bin/bash
mkdir -p gpurun_out/r2r
export LTXB_LIB=$PWD/mlx-video_b200/csrc/libltxb_wsdbg.so
for d in 8 24; do
for sh in "160x16384x4096 1" "160x16384x4096 2"; do
  echo "== LTXB_WS_DEBUG=$d"
  LTXB_WS_DEBUG=$d timeout 120 python scripts/gemm_small_m_trace.py $sh
done; done 2>&1 | tee gpurun_out/r2r/trace.txt
