#!/bin/bash
# kernel tests -> env A/B bench -> whole GPU suite.  Usage: bash scripts/gpu_ab_round.sh <tag> <settings...>
tag=$1; shift
out=gpurun_out/$tag; mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x > $out/pytest_kernels.log 2>&1; echo "kernel tests rc=$?"; tail -4 $out/pytest_kernels.log
timeout 200 python scripts/attn_sweep.py > $out/attn_sweep.txt 2>&1; cat $out/attn_sweep.txt
bash scripts/bench_env_ab.sh $tag "$@"
timeout 400 python -m pytest tests -m gpu -x -q > $out/pytest_all.log 2>&1; echo "all tests rc=$?"; tail -4 $out/pytest_all.log
