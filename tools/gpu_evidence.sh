#!/bin/bash
# One GPU-box call that refreshes the round's evidence: parity tests, bench lines, ncu launch list, ncu --set full captures.
# Usage (from the repo root, under gpurun): bash tools/gpu_evidence.sh <tag>
TAG=${1:-vX}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu_$TAG.log
python bench.py --steps 50 --warmup 10 > $O/bench_${TAG}_c2.json 2> $O/bench_${TAG}_c2.err; echo "bench rc=$?"; cut -c1-300 $O/bench_${TAG}_c2.json
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_${TAG}_ref.json 2>/dev/null
for w in c2fut c3 c5; do
  python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_${TAG}_$w.json 2> $O/bench_${TAG}_$w.err; cut -c1-160 $O/bench_${TAG}_$w.json
done
python tools/gpu_diag_model.py > $O/diag_$TAG.log 2>&1; grep -c "rel err" $O/diag_$TAG.log
python tools/gpu_probe_graph.py c2 > $O/probe_$TAG.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches_$TAG.csv \
  python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > $O/ncu_launches_$TAG.log 2>&1; echo "ncu launches rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gru_rw2 -s 12 -c 6 -f -o $O/ncu_full_rw_$TAG \
  python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > $O/ncu_full_$TAG.log 2>&1; echo "ncu full rw rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_p16 -s 93 -c 8 -f -o $O/ncu_full_gemm_$TAG \
  python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > $O/ncu_full_gemm_$TAG.log 2>&1; echo "ncu full gemm rc=$?"
ls -la $O | tail -30
