"""Timing of the k-means step at the C4 size: 10^6 latent vectors x 30 dims, 15 clusters, n_init 15
(the reference's defaults: vame/initialize_project/new.py:145-148).  Prints one JSON line:
  value       = latent vectors clustered per second for the whole fit (all restarts, seeding + Lloyd), device timed by wall clock
                around a synchronised call (the fit reads its convergence status once per iteration, so it is host-driven);
  roofline    = the Lloyd E+M pass (km_assign_kernel) timed alone with CUDA events: algorithmic bytes = 4*dim read + 4 written
                per point, against the measured HBM copy bandwidth;
  cpu_baseline = sklearn KMeans (the reference's call) on a bounded subset, scaled linearly in n (it is O(n) per iteration)."""
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from vame_b200 import _lib as L
from vame_b200.kmeans import DeviceKMeans


def main():
    n, d, k, n_init = 1_000_000, 30, 15, 15
    gen = torch.Generator(device="cuda").manual_seed(3)
    cent = torch.randn(k, d, device="cuda", generator=gen)
    X = cent[torch.randint(0, k, (n,), device="cuda", generator=gen)] + 1.5 * torch.randn(n, d, device="cuda", generator=gen)
    DeviceKMeans(k, n_init=1).fit(X[:50000])                      # warm-up (module load, attribute set-up)
    torch.cuda.synchronize()
    t0 = time.time()
    km = DeviceKMeans(k, random_state=42, n_init=n_init).fit(X)
    torch.cuda.synchronize()
    dt = time.time() - t0
    # the E+M pass alone
    lib = L.lib()
    ws = torch.empty(lib.vame_kmeans_workspace_bytes(n, d, k), dtype=torch.uint8, device="cuda")
    labels = torch.empty(n, dtype=torch.int32, device="cuda")
    cen = km.result.cluster_centers_
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        lib.vame_kmeans_assign(L.ptr(X), n, d, k, L.ptr(cen), L.ptr(labels), None, L.ptr(ws), ws.numel(), L.cur_stream())
    torch.cuda.synchronize()
    reps = 20
    e0.record()
    for _ in range(reps):
        lib.vame_kmeans_assign(L.ptr(X), n, d, k, L.ptr(cen), L.ptr(labels), None, L.ptr(ws), ws.numel(), L.cur_stream())
    e1.record()
    torch.cuda.synchronize()
    t_pass = e0.elapsed_time(e1) * 1e-3 / reps
    peak = 7700.0
    try:
        mp = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
        for key in ("hbm_gbps_sustained", "hbm_copy_gbps", "hbm_gbps", "hbm_gbps_burst"):
            if key in mp:
                peak = float(mp[key])
                break
    except Exception:
        pass
    ach = n * (4 * d + 4) / t_pass / 1e9
    # CPU: the reference's call on a subset
    from sklearn.cluster import KMeans
    ns = 100_000
    Xs = X[:ns].cpu().numpy()
    t0 = time.time()
    sk = KMeans(init="k-means++", n_clusters=k, random_state=42, n_init=3).fit(Xs)
    dts = time.time() - t0
    cpu_rate = ns * 3 / dts / n_init * 1.0                      # vectors/s for a 15-restart fit, scaled from 3 restarts
    out = {"metric": "latent_vectors_per_sec_kmeans_fit", "value": n / dt, "unit": "vectors/s", "n_gpus": 1, "seconds": dt,
           "config": {"workload": "k-means fit, n=%d dim=%d k=%d n_init=%d (k-means++ seeding + Lloyd), synthetic overlapping blobs" % (n, d, k, n_init)},
           "inertia": km.inertia_, "n_iter_best": km.n_iter_,
           "roofline": {"kernel": "km_assign_kernel (E-step pass)", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                        "frac": ach / peak, "us_per_launch": t_pass * 1e6,
                        "note": "fp64 distance accumulation (k*dim fp64 FMAs per point) keeps the trajectory independent of summation order; the pass is compute-bound on the fp64 pipe at k=15, dim=30, not HBM-bound"},
           "cpu_baseline": {"value": cpu_rate, "unit": "vectors/s", "cores": os.cpu_count(), "kind": "reference",
                            "sample": "sklearn KMeans(n_init=3) on the first %d vectors (%.2f s), scaled to n_init=%d" % (ns, dts, n_init)}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
