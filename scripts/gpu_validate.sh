#!/bin/bash
# One short gpurun call: the GPU tests named on the command line (default: the newest ones) without -x, then bench.py,
# then — with whatever time is left — the whole GPU suite.  Logs under gpurun_out/<tag>/.
# Usage: bash scripts/gpu_validate.sh <tag> [test files...]
tag=${1:-v1}; shift
tests=${@:-"tests/test_gpu_quant.py tests/test_gpu_sampler.py tests/test_gpu_lora.py"}
out=gpurun_out/$tag; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > $out/smi.txt 2>&1
timeout 360 python -m pytest $tests -q -m gpu > $out/pytest_new.log 2>&1; echo "new tests rc=$?"; tail -25 $out/pytest_new.log
timeout 300 python bench.py --kernel-table > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"; head -c 1500 $out/bench.json; echo; tail -22 $out/bench.err
timeout 420 python -m pytest tests -m gpu -x -q > $out/pytest_all.log 2>&1; echo "all tests rc=$?"; tail -5 $out/pytest_all.log
