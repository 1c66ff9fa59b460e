#!/bin/bash
# ncu capture of the attention kernel: plain run first (must exit 0), then one full-set capture.
out=gpurun_out/${1:-ncu_attn}; mkdir -p $out
args="${2:-1 5184 5184 32 128}"
python scripts/attn_one.py $args 5 > $out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attention_ -s 2 -c 1 -o $out/attn python scripts/attn_one.py $args 5 > $out/ncu.log 2>&1
echo rc=$?; tail -3 $out/plain.log; tail -5 $out/ncu.log
for a in "1 1280 1280 32 128" "1 1280 1024 32 128" "2 5184 5184 32 128" "1 14080 14080 32 128"; do python scripts/attn_one.py $a 20; done
