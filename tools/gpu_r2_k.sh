#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest tests -m gpu -q -x > $O/r2k_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -3 $O/r2k_pytest_all.log | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-300
for v in 1 0; do
for w in c2 c5; do
VAME_B200_EARLY_OPT=$v timeout 240 python bench.py --workload $w --steps 50 --warmup 10 --no-cpu-baseline --no-cudnn --no-train-epoch > $O/bench_r2k_${w}_eo$v.json 2> $O/bench_r2k_${w}_eo$v.err
python - <<PY
import json
try:
    d=json.loads(open('$O/bench_r2k_${w}_eo$v.json').read().splitlines()[-1])
    print('$w early_opt=$v', round(d['value']), 'w/s', round(d['ms_per_step'],4), 'ms; e2e', round(d['e2e']['value']), 'loss', d['loss_after'])
except Exception as e:
    print('FAILED', e); print(open('$O/bench_r2k_${w}_eo$v.err').read()[-800:])
PY
done
done
