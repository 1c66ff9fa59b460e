#!/bin/bash
mkdir -p gpurun_out/r2y
timeout 900 python -m pytest tests/test_gpu_quant.py -x -q > gpurun_out/r2y/pytest.log 2>&1
tail -5 gpurun_out/r2y/pytest.log
LTXB_BENCH_VARIANTS=small_m timeout 600 python scripts/gemm_small_m_bench.py 160x4096x4096 160x12288x4096 160x16384x4096 160x4096x16384 68x2048x2048 2>&1 | tee gpurun_out/r2y/sweep.txt
export LTXB_LIB=$PWD/mlx-video_b200/csrc/libltxb_wsdbg.so
echo "== no expansion arithmetic"
LTXB_WS_DEBUG=32 LTXB_BENCH_VARIANTS=small_m timeout 600 python scripts/gemm_small_m_bench.py 160x4096x4096 160x16384x4096 2>&1
