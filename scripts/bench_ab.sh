#!/bin/bash
# Same-box A/B of bench.py under different environments: bench_ab.sh "<env A>" "<env B>" [rounds] [extra bench args]
a="$1"; b="$2"; n=${3:-2}; shift 3
out=gpurun_out/bench_ab; mkdir -p $out
for i in $(seq 1 $n); do
  for v in A B; do
    if [ $v = A ]; then e="$a"; else e="$b"; fi
    env $e python bench.py --no-cpu-baseline "$@" > $out/$v$i.json 2> $out/$v$i.err
    python -c "
import json,sys
d=json.load(open('$out/$v$i.json')); print('$v$i', '[$e]', round(d['ms_per_step'],3), 'ms', round(d['value']), 'tok/s  e2e', round(d['e2e']['value']), d['clocks']['sm_mhz'], 'MHz', d['clocks'].get('power_w_max'))
"
  done
done
