#!/bin/bash
# Back-to-back timing of attention build variants (scripts/build_variant.sh): usage attn_ab.sh <tag>...
out=gpurun_out/attn_ab; mkdir -p $out; : > $out/summary.txt
shapes=("1 1280 1280 32 128" "2 5184 5184 32 128" "1 14080 14080 32 128")
for v in "" $@; do
  echo "== variant '${v}'" | tee -a $out/summary.txt
  for s in "${shapes[@]}"; do
    if [ -z "$v" ]; then timeout 120 python scripts/attn_one.py $s 20; else LTXB_LIB=mlx-video_b200/csrc/libltxb_$v.so timeout 120 python scripts/attn_one.py $s 20; fi
  done 2>&1 | tee -a $out/summary.txt
done
