#!/bin/bash
# ncu launch list of the eager C2 step with the final defaults
O=gpurun_out
mkdir -p $O
B="--steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-train-epoch --no-cudnn"
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 600 --csv --log-file $O/launches_c2_v16.csv python bench.py --workload c2 $B > $O/ncu_launches_c2_v16.log 2>&1; echo "ncu launches c2 rc=$?"
