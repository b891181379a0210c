#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_train_fn.py -m gpu -q -k "trained_checkpoint" > $O/r2e_pytest_trained.log 2>&1; echo "pytest trained rc=$?"; tail -5 $O/r2e_pytest_trained.log | cut -c1-400
timeout 400 python bench.py --steps 50 --warmup 10 > $O/bench_r2e_c2.json 2> $O/bench_r2e_c2.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('$O/bench_r2e_c2.json').read().splitlines()[-1])
    print(round(d['value']), 'w/s', round(d['ms_per_step'],4), 'ms e2e', round(d['e2e']['value']), 'clocks', d['clocks'])
    print('epoch', d['e2e_train_epoch'])
    print('cudnn', d['cudnn_reference']['value'], 'cpu', d['cpu_baseline'])
    print(d['roofline']['us_per_launch'], d['roofline']['step_frac_of_sustained_peak'])
except Exception as e:
    print('FAILED', e); print(open('$O/bench_r2e_c2.err').read()[-1500:])
PY
