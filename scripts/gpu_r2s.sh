#!/bin/bash
mkdir -p gpurun_out/r2s
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm" > gpurun_out/r2s/pytest_gemm.log 2>&1
tail -5 gpurun_out/r2s/pytest_gemm.log
timeout 600 python scripts/gemm_small_m_bench.py 16x16384x4096 68x2048x2048 160x4096x4096 160x12288x4096 160x16384x4096 160x4096x16384 256x4096x4096 256x16384x4096 320x4096x4096 320x12288x4096 2>&1 | tee gpurun_out/r2s/sweep.txt
export LTXB_LIB=$PWD/mlx-video_b200/csrc/libltxb_wsdbg.so
for sh in "160x16384x4096 0" "160x16384x4096 1" "160x4096x4096 0" "160x4096x16384 0" "160x12288x4096 0"; do
  LTXB_WS_DEBUG=8 timeout 120 python scripts/gemm_small_m_trace.py $sh
done 2>&1 | grep -v "met  \|-> met" | tee gpurun_out/r2s/trace.txt
