#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest tests -m gpu -q > $O/r2f_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -3 $O/r2f_pytest_all.log | cut -c1-300
timeout 300 python bench.py --workload c4 --steps 3 --warmup 3 > $O/bench_r2f_c4.json 2> $O/bench_r2f_c4.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('$O/bench_r2f_c4.json').read().splitlines()[-1])
    print(round(d['value']), 'w/s', round(d['ms_per_step'],1), 'ms; e2e', round(d['e2e']['value']), 'alg TFLOP/s', round(d['roofline']['achieved'],1), 'frac', round(d['roofline']['frac'],3), 'err', d['max_rel_err_vs_oracle_512_windows'], 'cpu', d['cpu_baseline'], 'launches', d['gpu_launches'])
except Exception as e:
    print('FAILED', e); print(open('$O/bench_r2f_c4.err').read()[-1500:])
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r2f_c4.csv python bench.py --workload c4 --per-gpu-batch 100000 --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_launches_r2f_c4.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py $O/launches_r2f_c4.csv 2>/dev/null | head -20
