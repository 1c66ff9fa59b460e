#!/bin/bash
mkdir -p gpurun_out/r2ae
timeout 900 python -m pytest tests/test_gpu_quant.py tests/test_gpu_kernels.py -x -q -k "gemm or packed" > gpurun_out/r2ae/pytest.log 2>&1
tail -3 gpurun_out/r2ae/pytest.log
LTXB_BENCH_VARIANTS=small_m timeout 600 python scripts/gemm_small_m_bench.py 160x4096x4096 160x12288x4096 160x16384x4096 160x4096x16384 68x2048x2048 320x4096x4096 320x16384x4096 2>&1 | tee gpurun_out/r2ae/sweep.txt
export LTXB_LIB=$PWD/mlx-video_b200/csrc/libltxb_wsdbg.so
LTXB_WS_DEBUG=64 LTXB_BENCH_VARIANTS=none LTXB_BENCH_PACKED=1 timeout 600 python scripts/gemm_small_m_bench.py 160x16384x4096 > /tmp/o.txt 2>&1
grep "MMA loop" /tmp/o.txt | tail -2; grep "expanding warp" /tmp/o.txt | tail -2; grep "producer" /tmp/o.txt | tail -2
