"""Regression guard for the root cause documented in profiles/r2_ng2_sw_fault.md: `tcgen05.ld` writes its destination registers
asynchronously, and every build in which ptxas had to SPILL inside a warp that executes it faulted on the GPU (and passed under
compute-sanitizer).  The tensor-core kernels must therefore compile without a stack frame; the one known exception (the
two-group forward sweep without store warps: 32 bytes - one kernel-wide scalar saved in the prologue and re-read once per role
before its loop, the rest inside the MMA-issue role, which executes no tcgen05.ld; checked in the SASS) is pinned here so
that any growth shows up on the CPU, before a GPU box is spent on it."""
import os
import re
import shutil
import subprocess

import pytest

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vame_b200", "libvame_b200.so")
# demangled-name fragment -> allowed stack bytes
ALLOWED = {"gru_rw2_fwd_kernelILb0ELb0ELi2E": 32, "gru_rw2_fwd_kernelILb1ELb0ELi2E": 32}


def _resource_usage():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    if not os.path.exists(LIB):
        pytest.skip("libvame_b200.so not built")
    out = subprocess.run([exe, "-res-usage", LIB], capture_output=True, text=True, timeout=120).stdout
    usage, name = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        if name and "REG:" in line:
            usage[name] = {k: int(v) for k, v in re.findall(r"(REG|STACK|LOCAL):(\d+)", line)}
            name = None
    return usage


def test_tensor_core_kernels_have_no_stack_frame():
    usage = _resource_usage()
    hot = {n: u for n, u in usage.items() if re.search(r"gru_rw2_|gru_rw_|gru_rows_|gemm_p16_kernel", n)}
    assert len(hot) >= 20, sorted(usage)
    bad = {}
    for n, u in hot.items():
        allow = max([v for k, v in ALLOWED.items() if k in n] + [0])
        if u["STACK"] > allow or u["LOCAL"] > 0:
            bad[n] = u
    assert not bad, bad


def test_sweep_kernels_fit_the_register_file():
    """One CTA per SM: threads x registers must fit 64 K registers (setmaxnreg budgets are checked on the GPU by the sweeps
    running at all; this pins the compile-time maxima)."""
    usage = _resource_usage()
    for n, u in usage.items():
        if "gru_rw2_" in n or "gru_rows_" in n:
            assert u["REG"] <= 255
