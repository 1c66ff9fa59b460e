#!/bin/bash
# in-model A/B of the small-M GEMM kernel: the 160-token stand-in and the audio+video workload, new kernel vs LTXB_GEMM_SMALL_M=0
mkdir -p gpurun_out/r2q
for v in 256 0; do
  LTXB_GEMM_SMALL_M=$v timeout 600 python bench.py --workload shard160 --workloads none --no-cpu-baseline --no-parity --steps 10 --warmup 3 --kernel-table > gpurun_out/r2q/shard160_small$v.json 2> gpurun_out/r2q/shard160_small$v.err
  python -c "import json;d=json.load(open('gpurun_out/r2q/shard160_small$v.json'));print('shard160 small_m=$v', d['ms_per_step'], {k:round(x['ms'],2) for k,x in d.get('kernels',{}).items()})"
done
for v in 256 0; do
  LTXB_GEMM_SMALL_M=$v timeout 600 python bench.py --workload av --workloads none --no-cpu-baseline --no-parity --steps 4 --warmup 3 --kernel-table > gpurun_out/r2q/av_small$v.json 2> gpurun_out/r2q/av_small$v.err
  python -c "import json;d=json.load(open('gpurun_out/r2q/av_small$v.json'));print('av small_m=$v', d['ms_per_step'], {k:round(x['ms'],2) for k,x in d.get('kernels',{}).items()})"
done
tail -3 gpurun_out/r2q/*.err
