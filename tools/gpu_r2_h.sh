#!/bin/bash
TAG=${1:-v14}
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest tests -m gpu -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest all rc=$?"; tail -3 $O/pytest_gpu_$TAG.log | cut -c1-300
run() { # name, args...
  n=$1; shift
  timeout 240 python bench.py "$@" --no-cpu-baseline --no-cudnn > $O/bench_${TAG}_$n.json 2> $O/bench_${TAG}_$n.err
  python - <<PY
import json
try:
    d=json.loads(open('$O/bench_${TAG}_$n.json').read().splitlines()[-1])
    r=d['roofline']
    print('$n', round(d['value']), 'w/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value']), r.get('us_per_launch') and {k.split(' ')[0]: round(v,2) for k,v in r['us_per_launch'].items()}, 'frac', round(r.get('step_frac_of_sustained_peak', r['frac']),4), d.get('e2e_train_epoch') and round(d['e2e_train_epoch']['device_sampler']))
except Exception as e:
    print('$n FAILED', e); print(open('$O/bench_${TAG}_$n.err').read()[-600:])
PY
}
run c2 --workload c2 --steps 50 --warmup 10
run c2fut --workload c2fut --steps 30 --warmup 5 --no-train-epoch
run c5 --workload c5 --steps 30 --warmup 5 --no-train-epoch
run c3 --workload c3 --steps 30 --warmup 5 --no-train-epoch
run c4 --workload c4 --steps 3 --warmup 3
