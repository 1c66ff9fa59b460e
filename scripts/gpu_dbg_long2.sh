#!/bin/bash
out=gpurun_out/dbg_long2; mkdir -p $out
run() { tag=$1; shift; timeout 900 "$@" > $out/$tag.json 2> $out/$tag.err; echo "$tag rc=$?"; grep -E "Error|error|ltxb:|Invalid|at 0x" $out/$tag.err | head -4; head -c 200 $out/$tag.json; echo; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29513"
B="bench.py --gpus 2 --workload long --workloads none --steps 3 --no-cpu-baseline --no-parity --layers 2"
run graph_l2 $TR $B
run graph_l2_nocache $TR $B --no-cache-context
LTXB_PDL=0 run graph_l2_nopdl $TR $B
