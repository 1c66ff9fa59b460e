#!/bin/bash
# Build the working-tree csrc with extra nvcc flags into mlx-video_b200/csrc/libltxb_<tag>.so for A/B runs
# (LTXB_LIB=...).  Usage: scripts/build_variant.sh <tag> <flags...>
set -e
tag=$1; shift
d=/tmp/ltxb_variant_$tag
rm -rf $d && mkdir -p $d/mlx-video_b200/csrc $d/include
cp -r /root/repo/include/. $d/include/
cp /root/repo/mlx-video_b200/csrc/*.cu /root/repo/mlx-video_b200/csrc/*.cuh /root/repo/mlx-video_b200/csrc/Makefile $d/mlx-video_b200/csrc/
make -C $d/mlx-video_b200/csrc -j4 EXTRA_NVCCFLAGS="$*" > /dev/null 2>&1
cp $d/mlx-video_b200/csrc/libltxb.so /root/repo/mlx-video_b200/csrc/libltxb_$tag.so
echo built libltxb_$tag.so
