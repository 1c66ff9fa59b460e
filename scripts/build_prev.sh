#!/bin/bash
# Build the committed (HEAD) csrc into mlx-video_b200/csrc/libltxb_prev.so for A/B runs (LTXB_LIB=...).
set -e
rm -rf /tmp/old_build && mkdir -p /tmp/old_build/include
git archive HEAD mlx-video_b200/csrc include | tar -x -C /tmp/old_build
cd /tmp/old_build/mlx-video_b200/csrc && make -j4 > /dev/null 2>&1
cp libltxb.so /root/repo/mlx-video_b200/csrc/libltxb_prev.so
