#!/bin/bash
# final state of round 2: smoke, the plain bench line (with CPU baseline, cuDNN comparator, train()-epoch), the other workloads
TAG=${1:-v16}
O=gpurun_out
mkdir -p $O
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-300
timeout 300 python bench.py > $O/bench_${TAG}_c2.json 2> $O/bench_${TAG}_c2.err; echo "bench rc=$?"
run() { n=$1; shift
  timeout 200 python bench.py "$@" --no-cpu-baseline --no-cudnn --no-train-epoch > $O/bench_${TAG}_$n.json 2> $O/bench_${TAG}_$n.err; }
run c2fut --workload c2fut --steps 40 --warmup 8
run c5 --workload c5 --steps 40 --warmup 8
run c3 --workload c3 --steps 40 --warmup 8
timeout 200 python bench.py --workload c4 --no-cpu-baseline > $O/bench_${TAG}_c4.json 2> $O/bench_${TAG}_c4.err
python - <<PY
import json
for n in ("c2","c2fut","c5","c3","c4"):
    try:
        d=json.loads(open('$O/bench_${TAG}_%s.json' % n).read().splitlines()[-1])
        r=d['roofline']
        print(n, round(d['value']), 'w/s', round(d['ms_per_step'],4), 'ms e2e', round(d['e2e']['value']), r.get('us_per_launch') and {k.split(' ')[0][7:]: round(v,2) for k,v in r['us_per_launch'].items()}, round(r.get('step_frac_of_sustained_peak', r['frac']),4), d.get('e2e_train_epoch') and round(d['e2e_train_epoch']['device_sampler']), d.get('cudnn_reference') and round(d['cudnn_reference']['value']), d.get('cpu_baseline') and round(d['cpu_baseline']['value']))
    except Exception as e:
        print(n, 'FAILED', e)
PY
