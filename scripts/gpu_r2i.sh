#!/bin/bash
# GEMM contiguous stream-K: kernel tests + small-M sweep; then the whole GPU suite and the bench line.
tag=${1:-r2i}; out=gpurun_out/$tag; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k gemm > $out/pytest_gemm.log 2>&1; echo "gemm tests rc=$?"; tail -4 $out/pytest_gemm.log; grep -E "^E  " $out/pytest_gemm.log | head -10
timeout 300 python scripts/gemm_small_m.py 2>&1 | tee $out/gemm_small_m.txt
echo "--- whole GPU suite"
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest_all.log 2>&1; echo "all tests rc=$?"; tail -6 $out/pytest_all.log
echo "--- bench"
timeout 900 python bench.py --kernel-table > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open("$out/bench.json"))
    print("bench", round(d["ms_per_step"], 3), "ms/step", round(d["value"]), "tok/s e2e", round(d["e2e"]["value"]), "e2e ms", round(d["e2e"]["ms_per_step"],3), "exec TF", round(d["model_tflops"]),
          "gemm frac", round(d["roofline"]["frac"], 3), "attn TF", d.get("attention_tflops"), d["clocks"], d["launch_mode"], d.get("parity"))
    print("cpu", d.get("cpu_baseline"))
    for k, v in d.get("workloads", {}).items():
        print(k, {kk: (round(vv, 2) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("value", "ms_per_step", "model_tflops", "error", "bench_wall_s", "attention_tflops")}, v.get("e2e", {}).get("ms_per_step"))
except Exception as e:
    print("bench parse failed", e)
PY
grep -A22 "^--- distilled" $out/bench.err | head -40
