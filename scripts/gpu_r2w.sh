#!/bin/bash
mkdir -p gpurun_out/r2w
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_quant.py -x -q > gpurun_out/r2w/pytest.log 2>&1
tail -4 gpurun_out/r2w/pytest.log
LTXB_BENCH_VARIANTS=small_m timeout 600 python scripts/gemm_small_m_bench.py 160x4096x4096 160x12288x4096 160x16384x4096 160x4096x16384 68x2048x2048 2>&1 | tee gpurun_out/r2w/sweep.txt
for i in 1 2; do
for v in new old; do
  if [ $v = old ]; then export LTXB_LIB=$PWD/mlx-video_b200/csrc/libltxb_relcluster.so; else unset LTXB_LIB; fi
  timeout 600 python bench.py --workloads none --no-cpu-baseline --no-parity --steps 12 --warmup 3 > gpurun_out/r2w/bench_$v$i.json 2> gpurun_out/r2w/bench_$v$i.err
  python -c "import json;d=json.load(open('gpurun_out/r2w/bench_$v$i.json'));print('$v', d['ms_per_step'], d['clocks']['sm_mhz'])"
done; done
