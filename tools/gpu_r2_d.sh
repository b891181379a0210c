#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -k "not trained_checkpoint" > $O/r2d_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -25 $O/r2d_pytest_all.log | cut -c1-300
timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_r2d_c2.json 2> $O/bench_r2d_c2.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.load(open('$O/bench_r2d_c2.json'))
    print(round(d['value']), 'w/s e2e', round(d['e2e']['value']), 'clocks', d['clocks'])
    print('epoch', d['e2e_train_epoch'])
    print('cudnn', d['cudnn_reference'])
except Exception as e:
    print('FAILED', e); print(open('$O/bench_r2d_c2.err').read()[-1500:])
PY
