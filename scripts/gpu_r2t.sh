#!/bin/bash
mkdir -p gpurun_out/r2t
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm" > gpurun_out/r2t/pytest_gemm.log 2>&1
tail -5 gpurun_out/r2t/pytest_gemm.log
SH="160x4096x4096 160x12288x4096 160x16384x4096 160x4096x16384"
LTXB_BENCH_VARIANTS=small_m,small_m_s1,small_m_s2 timeout 600 python scripts/gemm_small_m_bench.py $SH 2>&1 | tee gpurun_out/r2t/sweep_2persm.txt
LTXB_GEMM_SMALL_M_PER_SM=1 LTXB_BENCH_VARIANTS=small_m,small_m_s1,small_m_s2 timeout 600 python scripts/gemm_small_m_bench.py $SH 2>&1 | tee gpurun_out/r2t/sweep_1persm.txt
for v in 1 0; do
  LTXB_GEMM_CONST_W=$v timeout 600 python bench.py --workload shard160 --workloads none --no-cpu-baseline --no-parity --steps 10 --warmup 3 --kernel-table > gpurun_out/r2t/shard160_constw$v.json 2> gpurun_out/r2t/shard160_constw$v.err
  python -c "import json;d=json.load(open('gpurun_out/r2t/shard160_constw$v.json'));print('shard160 const_w=$v', d['ms_per_step'], {k:round(x['ms'],2) for k,x in list(d.get('kernels',{}).items())[:4]})"
done
