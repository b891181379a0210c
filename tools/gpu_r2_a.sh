#!/bin/bash
# round-2 check A: parity of the NG = 2 sweeps + sweep timings at C2 / C5 (short time-outs: a protocol bug must not eat the GPU budget)
O=gpurun_out
mkdir -p $O
timeout 240 python -m pytest tests/test_gpu_rw.py -x -q > $O/r2a_pytest_rw.log 2>&1; echo "pytest rw rc=$?"; tail -5 $O/r2a_pytest_rw.log
for args in "wl=c2" "wl=c2 rw_ng=2" "wl=c5" "wl=c5 rw_sw=1" "wl=c5 rw_exp=32" "wl=c3"; do
  timeout 90 python tools/gpu_probe_sweeps.py $args 2>&1 | tail -1 | cut -c1-900
done > $O/r2a_probe.log 2>&1
cat $O/r2a_probe.log
timeout 300 python -m pytest tests -m gpu -x -q > $O/r2a_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -3 $O/r2a_pytest_all.log
