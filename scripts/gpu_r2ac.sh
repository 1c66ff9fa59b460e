#!/bin/bash
mkdir -p gpurun_out/r2ac
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2ac/pytest.log 2>&1
tail -5 gpurun_out/r2ac/pytest.log
for wl in shard160 av distilled; do
  timeout 600 python bench.py --workload $wl --workloads none --no-cpu-baseline --no-parity --steps 8 --warmup 3 --kernel-table > gpurun_out/r2ac/$wl.json 2> gpurun_out/r2ac/$wl.err
  python -c "import json;d=json.load(open('gpurun_out/r2ac/$wl.json'));print('$wl', d['ms_per_step'], {k:round(x['ms'],2) for k,x in list(d.get('kernels',{}).items())[:5]})"
done
