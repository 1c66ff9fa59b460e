#!/bin/bash
mkdir -p gpurun_out/r2u
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_quant.py -x -q -k "gemm or quant or packed or dequant" > gpurun_out/r2u/pytest.log 2>&1
tail -15 gpurun_out/r2u/pytest.log
LTXB_BENCH_VARIANTS=small_m timeout 600 python scripts/gemm_small_m_bench.py 16x16384x4096 68x2048x2048 160x4096x4096 160x12288x4096 160x16384x4096 160x4096x16384 256x4096x4096 2>&1 | tee gpurun_out/r2u/sweep.txt
