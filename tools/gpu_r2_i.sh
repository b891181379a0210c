#!/bin/bash
O=gpurun_out
mkdir -p $O
( time timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" ) > $O/r2i_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/r2i_smoke.log | cut -c1-300
( time timeout 600 python bench.py ) > $O/r2i_bench_default.json 2> $O/r2i_bench_default.err; echo "bench default rc=$?"; grep -c "" $O/r2i_bench_default.json; tail -3 $O/r2i_bench_default.err
python - <<PY
import json
d=json.loads(open('$O/r2i_bench_default.json').read().splitlines()[-1])
print({k: d[k] for k in ('metric','value','n_gpus','steps','warmup','ms_per_step','gpu_launches','clocks')})
print(d['e2e'], d['e2e_train_epoch'], d['cudnn_reference'] and d['cudnn_reference']['value'], d['cpu_baseline'] and d['cpu_baseline']['value'])
print({k: v for k, v in d['roofline'].items() if k not in ('note','peak_source')})
PY
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) 2>&1 | tail -5 | cut -c1-400
( time timeout 600 python bench.py --impl reference --workload c4 --steps 1 --warmup 1 ) 2>&1 | tail -5 | cut -c1-400
