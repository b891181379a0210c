#!/bin/bash
# BASELINE configs[4]: data-parallel train step, 512 windows per GPU, global batch 4096 on 8 GPUs (one short run, strict time-out)
O=gpurun_out
mkdir -p $O
N=${1:-8}
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29752 bench.py --gpus $N --workload c5 --steps 50 --warmup 10 > $O/bench_r2_c5_n$N.json 2> $O/bench_r2_c5_n$N.err
python - <<PY
import json
try:
    d=json.loads(open('$O/bench_r2_c5_n$N.json').read().splitlines()[-1])
    print('c5 N=$N', round(d['value']), 'w/s', round(d['ms_per_step'],4), 'ms e2e', round(d['e2e']['value']), d['config']['global_batch'], d['clocks'])
except Exception as e:
    print('FAILED', e); print(open('$O/bench_r2_c5_n$N.err').read()[-800:])
PY
