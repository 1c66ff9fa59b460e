#!/bin/bash
out=gpurun_out/r2ak; mkdir -p $out
LTXB_BENCH_VARIANTS=none LTXB_BENCH_PACKED=1 timeout 300 python scripts/gemm_small_m_bench.py 160x16384x4096 > $out/plain.txt 2>&1; cat $out/plain.txt
LTXB_BENCH_VARIANTS=none LTXB_BENCH_PACKED=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_small_m_packed -s 40 -c 1 -o $out/packed python scripts/gemm_small_m_bench.py 160x16384x4096 > $out/ncu.log 2>&1
echo "ncu rc=$?"
ncu -i $out/packed.ncu-rep --page raw --csv > $out/packed.raw.csv 2>/dev/null; wc -c $out/packed.raw.csv
