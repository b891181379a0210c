"""Throughput of the training-set preparation on a synthetic (F=24, N=1e6) float64 series: the device path of
vame_b200.create_training (z-score, IQR, outlier removal + per-frame interpolation, Savitzky-Golay) with the series resident in
HBM and end to end from host numpy arrays, next to (a) the numpy oracle (vectorised restatement) on the same input and (b) the
reference's literal per-entry Python loop (create_training.py:220-231) timed on a 20 000-frame sample."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from oracle import trainset_numpy as tn
from vame_b200 import create_training as ct


def literal_loop(X_z, cut):
    """the reference's outlier pass, as written (create_training.py:220-231)"""
    for i in range(X_z.shape[0]):
        for marker in range(X_z.shape[1]):
            if X_z[i, marker] > cut:
                X_z[i, marker] = np.nan
            elif X_z[i, marker] < -cut:
                X_z[i, marker] = np.nan
        nans = np.isnan(X_z[i, :])
        if nans.any():
            idx = np.arange(X_z.shape[1])
            X_z[i, nans] = np.interp(idx[nans], idx[~nans], X_z[i, ~nans])
    return X_z


def main(N=1_000_000, F=24):
    rng = np.random.RandomState(5)
    x = np.cumsum(rng.randn(F, N) * 0.05, axis=1) + rng.randn(F, 1)
    idx = rng.choice(F * N, size=F * N // 200, replace=False)
    x.reshape(-1)[idx] += rng.choice([-1.0, 1.0], size=idx.size) * rng.uniform(20, 60, size=idx.size)
    dev = torch.from_numpy(x).cuda()
    for _ in range(2):
        xz, info = ct.zscore_clean(dev, True, 4, True)
        out = ct.savgol_filter(xz, 5, 2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    xz, info = ct.zscore_clean(dev, True, 4, True)
    out = ct.savgol_filter(xz, 5, 2)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    t0 = time.perf_counter()
    tr, te, _ = ct.trainset_arrays([x], True, True, 4, True, 5, 2, 0.1)
    t_e2e = time.perf_counter() - t0
    t0 = time.perf_counter()
    tr0, te0, _ = tn.traindata([x], True, True, 4, True, 5, 2, 0.1)
    t_np = time.perf_counter() - t0
    err = float(np.abs(tr - tr0).max() / np.abs(tr0).max())
    ns = 20000
    xs = (x[:, :ns].T - x[:, :ns].mean()) / x[:, :ns].std()
    t0 = time.perf_counter()
    literal_loop(xs.copy(), 4 * tn.iqr(xs))
    t_loop = time.perf_counter() - t0
    # one pass = 8 B read + 8 B written per entry; the path makes ~7 passes over the series plus the radix sort
    print(json.dumps({"workload": "create_trainset arithmetic, F=%d N=%d float64, robust fixed + savgol(5,2)" % (F, N),
                      "device_ms": ms, "frames_per_s_device": N / ms * 1e3, "gbytes_per_s_one_pass_equiv": F * N * 16 / ms / 1e6,
                      "e2e_s_from_host_numpy": t_e2e, "numpy_oracle_s": t_np, "outliers": info["outliers"],
                      "reference_literal_loop_s_per_20000_frames": t_loop, "reference_literal_loop_s_extrapolated": t_loop * N / ns,
                      "max_rel_err_vs_oracle": err}))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000)
