#!/bin/bash
# Round-2 evidence in ONE GPU-box call: embedding bench, ncu --set full captures of the new kernels, ncu launch lists, sanitizer logs.
TAG=${1:-v13}
O=gpurun_out
mkdir -p $O
B="--steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-train-epoch --no-cudnn"
timeout 200 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_${TAG}_c4.json 2> $O/bench_${TAG}_c4.err; echo "c4 rc=$?"; tail -1 $O/bench_${TAG}_c4.json | cut -c1-160
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gru_rows -s 2 -c 2 -f -o $O/ncu_full_rows_$TAG python bench.py --workload c4 --per-gpu-batch 20000 --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_full_rows_$TAG.log 2>&1; echo "ncu rows rc=$?"
timeout 250 ncu --set full --clock-control none --import-source on -k regex:gru_rw2 -s 12 -c 6 -f -o $O/ncu_full_rw_c5_$TAG python bench.py --workload c5 $B > $O/ncu_full_rw_c5_$TAG.log 2>&1; echo "ncu rw c5 rc=$?"
timeout 250 ncu --set full --clock-control none --import-source on -k regex:gru_rw2 -s 12 -c 6 -f -o $O/ncu_full_rw_c2_$TAG python bench.py --workload c2 $B > $O/ncu_full_rw_c2_$TAG.log 2>&1; echo "ncu rw c2 rc=$?"
timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 600 --csv --log-file $O/launches_c2_$TAG.csv python bench.py --workload c2 $B > $O/ncu_launches_c2_$TAG.log 2>&1; echo "ncu launches c2 rc=$?"
timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 600 --csv --log-file $O/launches_c5_$TAG.csv python bench.py --workload c5 $B > $O/ncu_launches_c5_$TAG.log 2>&1; echo "ncu launches c5 rc=$?"
timeout 200 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_rw.py -x -q -k "256-200-3-12-30-1-True-1" > $O/sanitizer_memcheck_rw_$TAG.log 2>&1; echo "memcheck rc=$?"; grep "ERROR SUMMARY\|passed\|failed" $O/sanitizer_memcheck_rw_$TAG.log | tail -3
timeout 300 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_rw.py -x -q -k "256-200-3-12-30-1-True-1" > $O/sanitizer_racecheck_rw_$TAG.log 2>&1; echo "racecheck rc=$?"; grep "RACECHECK SUMMARY\|passed\|failed" $O/sanitizer_racecheck_rw_$TAG.log | tail -3
ls -la $O | grep $TAG | awk '{print $5, $9}'
