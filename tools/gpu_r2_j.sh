#!/bin/bash
O=gpurun_out
mkdir -p $O
for v in 1 4; do
VAME_B200_ROWS=$v timeout 200 python -m pytest tests -m gpu -q -k "embed" > $O/r2j_pytest_embed_$v.log 2>&1; echo "pytest embed rows=$v rc=$?"; tail -1 $O/r2j_pytest_embed_$v.log | cut -c1-200
timeout 300 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --opt rows=$v > $O/bench_r2j_c4_rows$v.json 2> $O/bench_r2j_c4_rows$v.err
python - <<PY
import json
try:
    d=json.loads(open('$O/bench_r2j_c4_rows$v.json').read().splitlines()[-1])
    print('rows=$v', round(d['value']), 'w/s', round(d['ms_per_step'],1), 'ms; e2e', round(d['e2e']['value']), 'alg TFLOP/s', round(d['roofline']['achieved'],1), 'frac', round(d['roofline']['frac'],3), 'err', d['max_rel_err_vs_oracle_512_windows'])
except Exception as e:
    print('FAILED', e); print(open('$O/bench_r2j_c4_rows$v.err').read()[-1500:])
PY
done
