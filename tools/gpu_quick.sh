#!/bin/bash
# quick GPU check: parity tests + bench lines of all workloads (no CPU baseline)
TAG=${1:-q}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu_$TAG.log
for w in c2 c2fut c3 c5; do
  python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_${TAG}_$w.json 2> $O/bench_${TAG}_$w.err
  python -c "
import json,sys
d=json.load(open('$O/bench_${TAG}_$w.json'))
print('$w', round(d['value']), 'windows/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value']), d['roofline']['us_per_launch'])"
done
