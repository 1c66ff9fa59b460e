#!/bin/bash
export LTXB_LIB=$PWD/mlx-video_b200/csrc/libltxb_wsdbg.so
for sh in "160x4096x4096 0" "160x4096x16384 0" "160x12288x4096 0" "160x16384x4096 0"; do
  LTXB_WS_DEBUG=8 timeout 120 python scripts/gemm_small_m_trace.py $sh
done 2>&1
