#!/bin/bash
# Same-box A/B of environment switches: each argument is an "ENV=VALUE[,ENV=VALUE]" setting ("base" = none); every
# setting is benched twice, interleaved.  Usage: bash scripts/bench_env_ab.sh <tag> base LTXB_PREFETCH=1 ...
tag=$1; shift
out=gpurun_out/$tag; mkdir -p $out
for round in 1 2; do
  for setting in "$@"; do
    envs=""; [ "$setting" != base ] && envs=$(echo $setting | tr ',' ' ')
    f=$out/bench_${setting//[=,]/_}_$round.json
    env $envs timeout 300 python bench.py --no-cpu-baseline > $f 2> $f.err || echo "bench failed: $setting"
    python - <<PY
import json
try:
    d = json.load(open("$f"))
    print("$setting round $round:", round(d["ms_per_step"], 3), "ms/step", d["clocks"]["sm_mhz"], "MHz", d["gpu_launches"], "launches")
except Exception as e:
    print("$setting round $round: no result", e)
PY
  done
done | tee $out/summary.txt
