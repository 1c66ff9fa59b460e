#!/bin/bash
export LTXB_LIB=$PWD/mlx-video_b200/csrc/libltxb_wsdbg.so
for d in 0 128 160; do
echo "== LTXB_WS_DEBUG=$d"
LTXB_WS_DEBUG=$d LTXB_BENCH_VARIANTS=none LTXB_BENCH_PACKED=1 timeout 600 python scripts/gemm_small_m_bench.py 160x16384x4096 2>&1 | tail -1
done
