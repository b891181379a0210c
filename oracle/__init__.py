"""TEST INFRASTRUCTURE ONLY — CPU oracle for the VAME RNN-VAE hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / the CPU
baseline — never as the thing that is measured or shipped.

Contents
--------
* ``gru_numpy.py``   — numpy (fp64/fp32) restatement of the arithmetic the
  reference delegates to PyTorch (GRU cell, BPTT, Linear, losses, AMSGrad).
* ``vame_oracle.py`` — torch-CPU functional port of the reference modules on
  the hot path (same ATen calls the reference makes; used for autograd
  gradients and as the timed CPU baseline, kind="port").
* ``ref_shim.py``    — imports the UNMODIFIED reference from /root/reference
  (dev container only; that path does not exist on the GPU box).
* ``gen_golden.py``  — runs the real reference and writes tests/golden/*.npz.

Parity status: the reference ships no tests / golden vectors ("parity
unpinned" by the reference itself); the oracle is pinned instead against
outputs of the reference's own code run in the dev container
(tests/golden/*.npz, produced by gen_golden.py, checked by
tests/test_oracle_pinned.py).
"""
