#!/bin/bash
O=gpurun_out
mkdir -p $O
N=${1:-2}
run() { # tag workload env...
  tag=$1; w=$2; shift; shift
  env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29742 bench.py --gpus $N --workload $w --steps 50 --warmup 10 > $O/bench_r2_${w}_n${N}_$tag.json 2> $O/bench_r2_${w}_n${N}_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open('$O/bench_r2_${w}_n${N}_$tag.json').read().splitlines()[-1])
    print('$w N=$N $tag', round(d['value']), 'w/s', round(d['ms_per_step'],4), 'ms e2e', round(d['e2e']['value']))
except Exception as e:
    print('$w $tag FAILED', e); print(open('$O/bench_r2_${w}_n${N}_$tag.err').read()[-500:])
PY
}
run ov1_cta4 c2 VAME_B200_GRAD_OVERLAP=1 NCCL_MAX_CTAS=4
run ov1_cta2 c2 VAME_B200_GRAD_OVERLAP=1 NCCL_MAX_CTAS=2
run ov0_cta4 c2 VAME_B200_GRAD_OVERLAP=0 NCCL_MAX_CTAS=4
run ov0 c2 VAME_B200_GRAD_OVERLAP=0
run ov1_cta4 c5 VAME_B200_GRAD_OVERLAP=1 NCCL_MAX_CTAS=4
