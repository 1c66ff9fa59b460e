#!/bin/bash
# Multi-GPU validation on N GPUs of one box: parity (scripts/sp_check.py under torchrun), then the bench line the driver's
# SCALE run will produce.  Usage (gpurun --gpus N): bash scripts/gpu_r2_multi.sh <tag> <N> [bench args...]
tag=${1:-m2}; n=${2:-2}; shift; shift; out=gpurun_out/$tag; mkdir -p $out
nvidia-smi --query-gpu=index,name,memory.used --format=csv > $out/smi.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$n --master-addr 127.0.0.1 --master-port 29511 scripts/sp_check.py > $out/sp_check_$n.log 2>&1; echo "sp_check rc=$?"; grep -E "^(ok|FAIL)" $out/sp_check_$n.log | head -30; grep -iE "error|Traceback" $out/sp_check_$n.log | head -5
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --kernel-table "$@" > $out/bench_$n.json 2> $out/bench_$n.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open("$out/bench_$n.json"))
    print("bench N=$n", round(d["ms_per_step"], 3), "ms/step", round(d["value"]), "tok/s e2e ms", round(d["e2e"]["ms_per_step"],3), "exec TF", round(d["model_tflops"]),
          d["config"]["parallelism"], d["launch_mode"], d.get("parallel_parity"))
    for k, v in d.get("workloads", {}).items():
        print(k, {kk: (round(vv, 2) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("value", "ms_per_step", "model_tflops", "error", "bench_wall_s")}, v.get("config", {}).get("parallelism"), v.get("parallel_parity"))
except Exception as e:
    print("bench parse failed", e)
PY
grep -A16 "^--- distilled" $out/bench_$n.err | head -24; grep -iE "Traceback|Error" $out/bench_$n.err | head -5; tail -3 $out/bench_$n.err
