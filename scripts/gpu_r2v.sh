#!/bin/bash
mkdir -p gpurun_out/r2v
timeout 900 python -m pytest tests/test_gpu_quant.py -x -q -k "packed" > gpurun_out/r2v/pytest.log 2>&1
tail -3 gpurun_out/r2v/pytest.log
LTXB_BENCH_VARIANTS=small_m timeout 600 python scripts/gemm_small_m_bench.py 160x4096x4096 160x16384x4096 160x4096x16384 2>&1 | tee gpurun_out/r2v/sweep.txt
