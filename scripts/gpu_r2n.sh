#!/bin/bash
# small-M GEMM: where the time goes (debug build: no token loads / no epilogue / no weight loads), then ncu --set full
mkdir -p gpurun_out/r2n
export LTXB_LIB=$PWD/mlx-video_b200/csrc/libltxb_wsdbg.so
SH="160x16384x4096 160x4096x4096 16x16384x4096 256x16384x4096 160x4096x16384"
for d in 0 1 2 3 4 6 7; do
  echo "== LTXB_WS_DEBUG=$d"
  LTXB_WS_DEBUG=$d LTXB_BENCH_VARIANTS=small_m,small_m_s1 timeout 300 python scripts/gemm_small_m_bench.py $SH 2>&1 | grep -v "^ws trace"
done | tee gpurun_out/r2n/debug_knobs.txt
LTXB_WS_DEBUG=8 LTXB_BENCH_VARIANTS=small_m timeout 300 python scripts/gemm_small_m_bench.py 160x16384x4096 160x4096x4096 2>&1 | sort | uniq -c | sort -rn | head -30 > gpurun_out/r2n/trace.txt
cat gpurun_out/r2n/trace.txt
unset LTXB_LIB
for sh in 160x16384x4096 160x4096x4096; do
  LTXB_BENCH_VARIANTS=small_m timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_small_m -s 10 -c 2 -o gpurun_out/r2n/ws_$sh python scripts/gemm_small_m_bench.py $sh > gpurun_out/r2n/ncu_$sh.log 2>&1
  echo "ncu rc=$?"
  ncu -i gpurun_out/r2n/ws_$sh.ncu-rep --page raw --csv > gpurun_out/r2n/ws_$sh.raw.csv 2>/dev/null
done
ls -la gpurun_out/r2n
