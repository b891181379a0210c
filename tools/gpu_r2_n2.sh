#!/bin/bash
# N-GPU checks: multi-rank parity, weak-scaling bench lines with / without the bucketed allreduce overlap (strict time-outs)
O=gpurun_out
mkdir -p $O
N=${1:-2}
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29741 tools/dp_check.py 2>$O/r2_dp_check_n$N.err | grep "^{" > $O/r2_dp_check_n$N.json; echo "dp_check rc=$?"; cat $O/r2_dp_check_n$N.json; tail -3 $O/r2_dp_check_n$N.err | cut -c1-300
for v in 1; do
for w in c2; do
  VAME_B200_GRAD_OVERLAP=$v timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29742 bench.py --gpus $N --workload $w --steps 50 --warmup 10 > $O/bench_r2_${w}_n${N}_ov$v.json 2> $O/bench_r2_${w}_n${N}_ov$v.err
  python - <<PY
import json
try:
    d=json.loads(open('$O/bench_r2_${w}_n${N}_ov$v.json').read().splitlines()[-1])
    print('$w N=$N overlap=$v', round(d['value']), 'w/s', round(d['ms_per_step'],4), 'ms e2e', round(d['e2e']['value']))
except Exception as e:
    print('FAILED', e); print(open('$O/bench_r2_${w}_n${N}_ov$v.err').read()[-800:])
PY
done
done
