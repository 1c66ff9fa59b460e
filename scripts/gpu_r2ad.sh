#!/bin/bash
SH="160x4096x4096 160x12288x4096 320x4096x4096 320x12288x4096 320x16384x4096 320x4096x16384 68x2048x2048 68x2048x8192 68x8192x2048 2x36864x4096"
for ps in 0 1 2; do
echo "== LTXB_GEMM_SMALL_M_PER_SM=$ps"
LTXB_GEMM_SMALL_M_PER_SM=$ps LTXB_BENCH_PACKED=0 LTXB_BENCH_VARIANTS=small_m,small_m_s1,small_m_s2,small_m_s4,small_m_s8 timeout 600 python scripts/gemm_small_m_bench.py $SH 2>&1
done
