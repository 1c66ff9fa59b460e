#!/bin/bash
mkdir -p gpurun_out/r2ag
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm" > gpurun_out/r2ag/pytest.log 2>&1
tail -3 gpurun_out/r2ag/pytest.log
timeout 300 python scripts/gemm_sweep.py 2>&1 | head -8
bash scripts/gpu_ab_lib.sh mlx-video_b200/csrc/libltxb_oldprod.so
