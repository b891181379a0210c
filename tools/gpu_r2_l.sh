#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 200 python -m pytest tests -m gpu -q -k "embed" > $O/r2l_pytest_embed.log 2>&1; echo "pytest embed rc=$?"; tail -1 $O/r2l_pytest_embed.log | cut -c1-200
VAME_B200_RW_SW=7 timeout 300 python -m pytest tests -m gpu -q -x > $O/r2l_pytest_sw7.log 2>&1; echo "pytest rw_sw=7 rc=$?"; tail -1 $O/r2l_pytest_sw7.log | cut -c1-200
run() { # tag workload opts...
  tag=$1; w=$2; shift; shift
  timeout 120 python bench.py --workload $w --steps 40 --warmup 8 --no-cpu-baseline --no-cudnn --no-train-epoch "$@" > $O/bench_r2l_${w}_$tag.json 2> $O/bench_r2l_${w}_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open('$O/bench_r2l_${w}_$tag.json').read().splitlines()[-1])
    print('$w $tag', round(d['value']), 'w/s', round(d['ms_per_step'],4), 'ms', {k.split(' ')[0][7:]: round(v,2) for k,v in d['roofline']['us_per_launch'].items()})
except Exception as e:
    print('$w $tag FAILED', e); print(open('$O/bench_r2l_${w}_$tag.err').read()[-400:])
PY
}
run base c2
run sw7 c2 --opt rw_sw=7
run side20 c2 --opt side_sms=20
run side32 c2 --opt side_sms=32
run side74 c2 --opt side_sms=74
run base c5
run side20 c5 --opt side_sms=20
run side74 c5 --opt side_sms=74
timeout 200 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c4', round(d['value']), round(d['ms_per_step'],1), d['max_rel_err_vs_oracle_512_windows'])"
