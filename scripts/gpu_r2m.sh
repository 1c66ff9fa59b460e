#!/bin/bash
# small-M GEMM kernel: tests, then the per-launch sweep (new kernel / old kernel / cuBLAS)
set -x
mkdir -p gpurun_out/r2m
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm" > gpurun_out/r2m/pytest_gemm.log 2>&1
tail -15 gpurun_out/r2m/pytest_gemm.log
timeout 600 python scripts/gemm_small_m_bench.py > gpurun_out/r2m/sweep_new.txt 2>&1
LTXB_GEMM_SMALL_M=0 timeout 600 python scripts/gemm_small_m_bench.py 160x4096x4096 160x12288x4096 160x16384x4096 160x4096x16384 68x2048x2048 > gpurun_out/r2m/sweep_old.txt 2>&1
cat gpurun_out/r2m/sweep_new.txt
cat gpurun_out/r2m/sweep_old.txt
