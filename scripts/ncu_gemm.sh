#!/bin/bash
# ncu capture of the GEMM kernel: plain run first (must exit 0), then one full-set capture.
out=gpurun_out/${1:-ncu1}; mkdir -p $out
args="${2:-5184 4096 4096 1 256}"
python scripts/gemm_one.py $args 5 > $out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 2 -c 2 -o $out/gemm python scripts/gemm_one.py $args 5 > $out/ncu.log 2>&1
echo rc=$?; tail -3 $out/plain.log; tail -5 $out/ncu.log
