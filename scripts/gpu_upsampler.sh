#!/bin/bash
tag=${1:-u1}; out=gpurun_out/$tag; mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_upsampler.py -q -m gpu > $out/pytest_upsampler.log 2>&1; echo "upsampler tests rc=$?"; tail -30 $out/pytest_upsampler.log
timeout 200 python scripts/upsampler_bench.py 2>&1 | tail -3 | tee $out/upsampler_bench.json
