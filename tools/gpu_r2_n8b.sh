#!/bin/bash
O=gpurun_out
mkdir -p $O
N=${1:-8}
for v in 0 1; do
VAME_B200_GRAD_OVERLAP=$v timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29753 bench.py --gpus $N --workload c2 --steps 50 --warmup 10 > $O/bench_r2_c2_n${N}_ov$v.json 2> $O/bench_r2_c2_n${N}_ov$v.err
python - <<PY
import json
try:
    d=json.loads(open('$O/bench_r2_c2_n${N}_ov$v.json').read().splitlines()[-1])
    print('c2 N=$N overlap=$v', round(d['value']), 'w/s', round(d['ms_per_step'],4), 'ms e2e', round(d['e2e']['value']))
except Exception as e:
    print('FAILED', e); print(open('$O/bench_r2_c2_n${N}_ov$v.err').read()[-800:])
PY
done
