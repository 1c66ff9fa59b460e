#!/bin/bash
# Timing experiment: the hot GEMM shapes with the default library and with variant builds named on the command line.
# (The skip-A / skip-W timing macros used for profiles/r1e/gemm_operand_delivery_exp.txt live in
#  profiles/r1e/quad_multicast_experiment.patch, not in the product sources.)
# Usage: bash scripts/gemm_exp.sh <tag> <variant> [<variant> ...]   (variants are libltxb_<variant>.so; "base" = libltxb.so)
tag=$1; shift
out=gpurun_out/$tag; mkdir -p $out
for v in base "$@"; do
  lib=mlx-video_b200/csrc/libltxb_$v.so; [ $v = base ] && lib=mlx-video_b200/csrc/libltxb.so
  for shape in "1280 12288 4096" "1280 16384 4096" "1280 4096 16384" "1280 4096 4096" "5184 16384 4096"; do
    echo -n "$v " ; LTXB_LIB=$lib timeout 120 python scripts/gemm_one.py $shape -1 0 30 0 2>&1 | tail -1
  done
done | tee $out/gemm_exp.txt
