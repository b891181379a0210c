#!/bin/bash
# bench lines of all workloads (no CPU baseline) + variants
TAG=${1:-r2c}
O=gpurun_out
mkdir -p $O
run() { # name, args...
  n=$1; shift
  timeout 240 python bench.py "$@" --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_${TAG}_$n.json 2> $O/bench_${TAG}_$n.err
  python - <<PY
import json
try:
    d=json.load(open('$O/bench_${TAG}_$n.json'))
    print('$n', round(d['value']), 'windows/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value']), {k.split(' ')[0]: round(v,2) for k,v in d['roofline']['us_per_launch'].items()}, 'step_frac_sus', round(d['roofline']['step_frac_of_sustained_peak'],4))
except Exception as e:
    print('$n FAILED', e); print(open('$O/bench_${TAG}_$n.err').read()[-600:])
PY
}
run c2 --workload c2
run c2_ng2 --workload c2 --opt rw_ng=2
run c2fut --workload c2fut
run c5 --workload c5
run c5_ng1 --workload c5 --opt rw_ng=1
run c3 --workload c3
run c3_ng1 --workload c3 --opt rw_ng=1
