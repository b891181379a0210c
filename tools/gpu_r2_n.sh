#!/bin/bash
# final defaults: full GPU test suite; experiment: per-direction split of the side-GEMM cap
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests -m gpu -q -x > $O/r2n_pytest_all.log 2>&1; echo "pytest rc=$?"; tail -1 $O/r2n_pytest_all.log | cut -c1-200
run() { tag=$1; w=$2; shift; shift
  timeout 120 python bench.py --workload $w --steps 40 --warmup 8 --no-cpu-baseline --no-cudnn --no-train-epoch "$@" > $O/bench_r2n_${w}_$tag.json 2> $O/bench_r2n_${w}_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open('$O/bench_r2n_${w}_$tag.json').read().splitlines()[-1])
    print('$w $tag', round(d['value']), 'w/s', round(d['ms_per_step'],4), 'ms e2e', round(d['e2e']['value']))
except Exception as e:
    print('$w $tag FAILED', e); print(open('$O/bench_r2n_${w}_$tag.err').read()[-400:])
PY
}
run split1 c2 --opt side_split=1
run split2 c2
run split1 c5 --opt side_split=1
run split2 c5
run split1 c3 --opt side_split=1
run split1 c2fut --opt side_split=1
