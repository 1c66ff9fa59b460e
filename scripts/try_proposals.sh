#!/bin/bash
# First GPU call of the next round: try the two compile-checked, never-run proposals of profiles/r1e/ in a scratch copy of
# the tree (the product tree is not touched), each against the current build:
#   1. attention_pair64 (double-buffered 64-key score tiles)      LTXB_ATTN_S64=1
#   2. implicit-GEMM convolutions for the latent upsampler        LTXB_CONV_IMPLICIT=1
# Usage (on the GPU box, via gpurun):  bash scripts/try_proposals.sh <tag>
tag=${1:-p1}; root=$(pwd); out=$root/gpurun_out/$tag; mkdir -p $out
w=/tmp/ltxb_proposals; rm -rf $w; mkdir -p $w; cp -r $root/. $w/ 2>/dev/null; cd $w
patch -p1 -s < profiles/r1e/attention_pair64_wiring.patch || { echo "attention patch does not apply"; exit 1; }
patch -p1 -s < profiles/r1e/implicit_conv_proposal.patch || { echo "implicit-conv patch does not apply"; exit 1; }
make -C mlx-video_b200/csrc -j8 > $out/build.log 2>&1 || { tail -20 $out/build.log; echo "build failed"; exit 1; }
echo "--- patched tree, default paths (must stay green)"
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_upsampler.py -q -m gpu -x > $out/pytest_default.log 2>&1; echo "rc=$?"; tail -3 $out/pytest_default.log
echo "--- attention_pair64"
LTXB_ATTN_S64=1 timeout 200 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k attention > $out/pytest_s64.log 2>&1; echo "rc=$?"; tail -6 $out/pytest_s64.log
LTXB_ATTN_S64=1 timeout 200 python scripts/attn_sweep.py 2>&1 | tee $out/attn_sweep_s64.txt
timeout 200 python scripts/attn_sweep.py 2>&1 | tee $out/attn_sweep_base.txt
echo "--- implicit-GEMM convolutions"
LTXB_CONV_IMPLICIT=1 timeout 200 python -m pytest tests/test_gpu_upsampler.py -q -m gpu > $out/pytest_iconv.log 2>&1; echo "rc=$?"; tail -6 $out/pytest_iconv.log
LTXB_CONV_IMPLICIT=1 timeout 100 python scripts/upsampler_bench.py 2>&1 | tail -1 | tee $out/upsampler_iconv.json
timeout 100 python scripts/upsampler_bench.py 2>&1 | tail -1 | tee $out/upsampler_base.json
echo "--- bench A/B (attention variant inside the step)"
for v in 0 1 0 1; do
  LTXB_ATTN_S64=$v timeout 200 python bench.py --no-cpu-baseline > $out/bench_s64_$v.json 2> /dev/null
  python -c "import json; d=json.load(open('$out/bench_s64_$v.json')); print('LTXB_ATTN_S64=$v', round(d['ms_per_step'],3), 'ms/step')" 2>/dev/null || echo "LTXB_ATTN_S64=$v: bench failed"
done
