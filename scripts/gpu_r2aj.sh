#!/bin/bash
mkdir -p gpurun_out/r2aj
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -x -q -k "attention or block or model or graph" > gpurun_out/r2aj/pytest.log 2>&1
tail -3 gpurun_out/r2aj/pytest.log
for f in 1 0; do
  echo "== LTXB_ATTN_FUSED_COMBINE=$f"
  LTXB_ATTN_FUSED_COMBINE=$f timeout 300 python scripts/attn_sweep.py 2>&1 | cut -c1-140
done
for f in 1 0 1 0; do
  LTXB_ATTN_FUSED_COMBINE=$f timeout 600 python bench.py --workloads none --no-cpu-baseline --no-parity --steps 12 --warmup 3 > gpurun_out/r2aj/bench_f$f.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/r2aj/bench_f$f.json'));print('fused=$f', round(d['ms_per_step'],3), d['clocks']['sm_mhz'], {k:round(x['ms'],3) for k,x in list(d['kernels'].items())[:2]})"
done
