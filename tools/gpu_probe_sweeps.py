"""Probe: time the encoder-layer-1 forward / backward sweeps alone (vame_debug_gru_sweep) for option variants given as
name=value pairs on the command line, e.g.  python tools/gpu_probe_sweeps.py rw2=1 rw_exp=3  (wl=c5 selects the workload)"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from bench import WORKLOADS
    from oracle import vame_oracle as vo
    from vame_b200.engine import Engine
    from vame_b200 import _lib as L
    lib = L.lib()
    wl = "c2"
    for kv in sys.argv[1:]:
        k, v = kv.split("=")
        if k == "wl":
            wl = v
            continue
        lib.vame_set_option(k.encode(), int(v))
    F, T, Z, H, fut, S, B = WORKLOADS[wl]
    torch.manual_seed(19)
    port = vo.RefPort(2 * T, Z, F, fut, S, hidden=H)
    eng = Engine(F, T, Z, H, H, H, fut, S, False, device="cuda")
    eng.load_state_dict(port.state_dict())
    x, xf, eps = vo.synthetic_batch(B, T, F, max(S, 1), Z)
    cfg = eng.loss_cfg(kmeans_loss=Z, kmeans_lambda=0.1, bsize=B, beta=1.0, kl_weight=1.0)
    eng.forward(x.cuda(), eps.cuda(), save=True, want=())
    eng.loss(cfg, None, want_grads=True)
    eng.backward(cfg)
    torch.cuda.synchronize()
    ws = eng.workspace(B, True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = []
    for which in (0, 1):
        dbg = torch.zeros(64, dtype=torch.int64, device="cuda")
        lib.vame_set_debug_buffer(ctypes.c_void_p(dbg.data_ptr()))
        for _ in range(3):
            lib.vame_debug_gru_sweep(ctypes.byref(eng.dims), B, which, L.ptr(eng.flat), L.ptr(eng.packed), L.ptr(ws), ws.numel(), L.cur_stream())
        torch.cuda.synchronize()
        lib.vame_set_debug_buffer(None)
        st = dbg.cpu().tolist()
        e0.record()
        for _ in range(20):
            lib.vame_debug_gru_sweep(ctypes.byref(eng.dims), B, which, L.ptr(eng.flat), L.ptr(eng.packed), L.ptr(ws), ws.numel(), L.cur_stream())
        e1.record()
        torch.cuda.synchronize()
        out.append("%s %.2f us/step stamps %s" % ("fwd" if which == 0 else "bwd", e0.elapsed_time(e1) * 1e3 / (20 * T),
                                                   [(st[i] - st[0]) if st[i] else None for i in range(12)]))
        if any(st[16:48]):      # per-CTA stamps of one cluster (rw_exp & 32)
            n = 4 if which == 0 else 8
            out.append("per-CTA %s" % [[(st[16 + n * c + i] - st[0]) if st[16 + n * c + i] else None for i in range(n)] for c in range(4)])
    print(" ".join(sys.argv[1:]) or "defaults", "|", " | ".join(out), "| timeouts", lib.vame_get_option(b"rw_timeouts"), flush=True)


if __name__ == "__main__":
    main()
