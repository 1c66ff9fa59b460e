#!/bin/bash
# folded peer barriers: parity, then same-box A/B against the stand-alone barrier launches.  Usage (gpurun --gpus N): bash scripts/gpu_r2_fold.sh <N>
n=${1:-2}; out=gpurun_out/r2fold$n; mkdir -p $out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$n --master-addr 127.0.0.1 --master-port 29511 scripts/sp_check.py > $out/sp_check_fold.log 2>&1; echo "sp_check (folded) rc=$?"; grep -E "^(ok|FAIL)" $out/sp_check_fold.log | head -30; grep -iE "error|Traceback|watchdog" $out/sp_check_fold.log | head -5
for f in 1 0 1 0; do
  LTXB_FOLD_BARRIERS=$f timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --workloads none --no-cpu-baseline --steps 12 --warmup 3 --kernel-table > $out/bench_fold$f.json 2> $out/bench_fold$f.err; echo "bench fold=$f rc=$?"
  python -c "import json;d=json.load(open('$out/bench_fold$f.json'));print('fold=$f', round(d['ms_per_step'],3), d.get('parallel_parity'), {k:(round(x['ms'],2),x.get('launches')) for k,x in list(d.get('kernels',{}).items())[:7]})"
done
grep -iE "Traceback|Error|watchdog" $out/*.err | head
