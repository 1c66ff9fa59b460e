#!/bin/bash
# Round 2, first GPU call: validate the double-buffered 64-key attention kernel (default) against the single-buffered one
# (LTXB_ATTN_S64=0), sweep both (+ exp2-emulation variants), trace one CTA, then the whole GPU suite and the bench line.
# Usage: bash scripts/gpu_r2a.sh <tag>
tag=${1:-r2a}; out=gpurun_out/$tag; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > $out/smi.txt 2>&1
echo "--- attention kernel tests, S64 (default)"
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k attention > $out/pytest_attn_s64.log 2>&1; rc=$?; echo "rc=$rc"; tail -5 $out/pytest_attn_s64.log
if [ $rc -ne 0 ]; then echo "S64 kernel FAILED its tests: the rest of this call runs with LTXB_ATTN_S64=0"; export LTXB_ATTN_S64=0; grep -m5 -B2 -A12 "Error\|assert" $out/pytest_attn_s64.log | head -60; fi
echo "--- sweeps"
LTXB_ATTN_S64=0 timeout 200 python scripts/attn_sweep.py 2>&1 | tee $out/attn_sweep_s128.txt
if [ "$LTXB_ATTN_S64" != "0" ]; then
  timeout 200 python scripts/attn_sweep.py 2>&1 | tee $out/attn_sweep_s64.txt
  for v in emu0 emu1 emu3 emu4; do
    [ -f mlx-video_b200/csrc/libltxb_$v.so ] && LTXB_LIB=$PWD/mlx-video_b200/csrc/libltxb_$v.so ATTN_SWEEP_NO_SDPA=1 timeout 120 python scripts/attn_sweep.py 2>&1 | tee $out/attn_sweep_s64_$v.txt
  done
  for shape in "1 5184 5184 32 128" "1 1280 1280 32 128"; do
    LTXB_LIB=$PWD/mlx-video_b200/csrc/libltxb_trace.so timeout 60 python scripts/attn_trace.py $shape > "$out/attn_trace_$(echo $shape | tr ' ' '_').txt" 2>&1
  done
fi
echo "--- whole GPU suite"
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest_all.log 2>&1; echo "all tests rc=$?"; tail -15 $out/pytest_all.log
echo "--- bench"
timeout 900 python bench.py --kernel-table > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open("$out/bench.json"))
    print("bench", round(d["ms_per_step"], 3), "ms/step", round(d["value"]), "tok/s e2e", round(d["e2e"]["value"]), "exec TF", round(d["model_tflops"]),
          "gemm frac", round(d["roofline"]["frac"], 3), "attn TF", d.get("attention_tflops"), d["clocks"], d["launch_mode"], d.get("parity"))
    print("cpu", d.get("cpu_baseline"))
    for k, v in d.get("workloads", {}).items():
        print(k, {kk: (round(vv, 2) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("value", "ms_per_step", "model_tflops", "error", "bench_wall_s", "attention_tflops")}, v.get("e2e", {}).get("ms_per_step"))
except Exception as e:
    print("bench parse failed", e)
PY
tail -40 $out/bench.err
