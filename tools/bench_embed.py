"""Throughput + parity spot-check of the sliding-window embedding (BASELINE configs[3]: 1e6-frame synthetic series,
F=24, T=30, Z=30): windows/s through the public embed path with the series resident in HBM, and end-to-end from a host
numpy array (the call embedd_latent_vectors makes).  2000 random windows are checked against the CPU oracle."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from oracle import vame_oracle as vo
from vame_b200.engine import Engine


def main(n_frames=1_000_000, chunk=8192):
    F, T, Z, H = 24, 30, 30, 256
    torch.manual_seed(19)
    port = vo.RefPort(2 * T, Z, F, True, 15, hidden=H)
    eng = Engine(F, T, Z, H, H, H, True, 15, False, device="cuda")
    eng.load_state_dict(port.state_dict())
    rng = np.random.default_rng(5)
    series = rng.standard_normal((n_frames, F)).astype(np.float32)      # z-scored pose-like input
    dev = torch.from_numpy(series).cuda()
    lat = eng.embed(dev, chunk=chunk)                                    # warm-up (allocates the workspace)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lat = eng.embed(dev, chunk=chunk)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    n_win = lat.shape[0]
    t0 = time.perf_counter()
    host = torch.from_numpy(series).pin_memory()
    lat2 = eng.embed(host.cuda(non_blocking=True), chunk=chunk).cpu()
    t_e2e = time.perf_counter() - t0
    # parity: random windows vs the CPU oracle
    idx = np.sort(rng.choice(n_win, size=2000, replace=False))
    xw = torch.from_numpy(np.stack([series[i:i + T] for i in idx]))
    with torch.no_grad():
        ref = port.lmbda(port.encode(xw), None)[1]
    err = float((lat[idx].cpu() - ref).abs().max() / ref.abs().max())
    flops = 96.71e6 * n_win
    out = {"workload": "embedd_latent_vectors, %d frames, F=24 T=30 Z=30 H=256, chunk %d" % (n_frames, chunk), "windows": int(n_win),
           "ms": ms, "windows_per_s": n_win / ms * 1e3, "algorithmic_tflops": flops / (ms * 1e-3) / 1e12,
           "e2e_s_from_host_numpy": t_e2e, "e2e_windows_per_s": n_win / t_e2e, "max_rel_err_vs_oracle_2000_windows": err,
           "chunk_invariance": bool(torch.equal(lat2, lat.cpu()))}
    print(json.dumps(out))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, int(sys.argv[2]) if len(sys.argv) > 2 else 8192)
