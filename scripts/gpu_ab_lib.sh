#!/bin/bash
# same-box A/B of the product library against a variant build (LTXB_LIB): alternating runs of the headline bench + kernel table.
# Usage: bash scripts/gpu_ab_lib.sh <variant .so> [bench args...]
v=$1; shift; out=gpurun_out/ab_$(basename $v .so); mkdir -p $out
for i in 1 2; do
  for which in base var; do
    if [ $which = var ]; then export LTXB_LIB=$PWD/$v; else unset LTXB_LIB; fi
    timeout 600 python bench.py --workloads none --no-cpu-baseline --no-parity --steps 12 --warmup 3 --kernel-table "$@" > $out/${which}$i.json 2> $out/${which}$i.err
    python -c "import json;d=json.load(open('$out/${which}$i.json'));print('$which', round(d['ms_per_step'],3), d['clocks']['sm_mhz'], {k:round(x['ms'],3) for k,x in list(d['kernels'].items())[:6]})"
  done
done
