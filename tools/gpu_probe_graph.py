"""Probe: per-phase timing of the train step in eager / CUDA-graph mode with PDL on/off, each variant in its own
subprocess under a timeout (a hang in one variant must not take the others down)."""
import faulthandler
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(mode, pdl, workload, streams=1, persistent=1, flags=1):
    faulthandler.dump_traceback_later(100, exit=True)
    import torch
    from bench import WORKLOADS
    from oracle import vame_oracle as vo
    from vame_b200.engine import Engine, TrainStep
    from vame_b200 import _lib as L
    F, T, Z, H, fut, S, B = WORKLOADS[workload]
    lib = L.lib()
    lib.vame_set_option(b"pdl", int(pdl))
    lib.vame_set_option(b"streams", int(streams))
    lib.vame_set_option(b"persistent", int(persistent))
    lib.vame_set_option(b"flags", int(flags))
    torch.manual_seed(19)
    port = vo.RefPort(2 * T, Z, F, fut, S, hidden=H)
    eng = Engine(F, T, Z, H, H, H, fut, S, False, device="cuda")
    eng.load_state_dict(port.state_dict())
    x, xf, eps = vo.synthetic_batch(B, T, F, max(S, 1), Z)
    cfg = eng.loss_cfg(kmeans_loss=Z, kmeans_lambda=0.1, bsize=B, beta=1.0, kl_weight=1.0)
    eng.set_hyper(lr=5e-4, kl_weight=1.0, beta=1.0, kmeans_lambda=0.1)
    ts = TrainStep(eng, B, cfg, world=1, use_graph=(mode == "graph"))
    print("[%s pdl=%d] capture..." % (mode, pdl), flush=True)
    ok = ts.capture()
    print("[%s pdl=%d] captured=%s err=%s" % (mode, pdl, ok, getattr(ts, "capture_error", None)), flush=True)
    ts.load(x.cuda(), xf[:, :S].cuda() if fut else None, eps.cuda())
    for i in range(3):
        ts.run()
        torch.cuda.synchronize()
        print("[%s pdl=%d] warm step %d ok" % (mode, pdl, i), flush=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        ts.run()
    e1.record()
    torch.cuda.synchronize()
    print("[%s pdl=%d] RESULT %s ms/step %.3f  windows/s %.0f  rw_timeouts=%d info=%s" % (mode, pdl, workload, e0.elapsed_time(e1) / n, B * n / e0.elapsed_time(e1) * 1e3,
          lib.vame_get_option(b"rw_timeouts"), [lib.vame_get_option(b"rw_timeout_info%d" % i) for i in range(6)]), flush=True)
    if mode == "eager":
        # per-phase breakdown with events
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        tot = [0.0] * 4
        for _ in range(10):
            ev[0].record()
            eng.forward(ts.x, ts.eps, save=True, want=(), ensure_packed=False)
            ev[1].record()
            eng.loss(cfg, ts.fut if fut else None, want_grads=True, use_hyper=True, out=ts.losses)
            ev[2].record()
            eng.backward(cfg, use_hyper=True)
            ev[3].record()
            eng.adam_step(use_hyper=True, repack=True)
            ev[4].record()
            torch.cuda.synchronize()
            for i in range(4):
                tot[i] += ev[i].elapsed_time(ev[i + 1]) / 10
        print("[%s pdl=%d] phases ms: forward %.3f loss %.3f backward %.3f adam+repack %.3f" % ((mode, pdl) + tuple(tot)), flush=True)
        # section timeline of one eager step (cudaEvents at section boundaries on the main stream)
        import ctypes as _ct
        lib.vame_debug_timeline(1)
        eng.forward(ts.x, ts.eps, save=True, want=(), ensure_packed=False)
        eng.loss(cfg, ts.fut if fut else None, want_grads=True, use_hyper=True, out=ts.losses)
        eng.backward(cfg, use_hyper=True)
        names = (_ct.c_char_p * 64)()
        msv = (_ct.c_float * 64)()
        nsec = lib.vame_debug_timeline_read(names, msv, 64)
        lib.vame_debug_timeline(0)
        print("[%s pdl=%d] SECTIONS (us since fwd:start) " % (mode, pdl) + " | ".join("%s @%.0f" % (names[i].decode(), msv[i] * 1e3) for i in range(nsec)), flush=True)
        # host-side launch cost
        t0 = time.perf_counter()
        for _ in range(5):
            ts._phase1(); ts._phase2()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        print("[%s pdl=%d] host enqueue time per step %.3f ms" % (mode, pdl, (t1 - t0) / 5 * 1e3), flush=True)
        # in-kernel timeline of one forward step kernel (CTA 0): globaltimer stamps
        import ctypes
        dbg = torch.zeros(16, dtype=torch.int64, device="cuda")
        lib.vame_set_debug_buffer(ctypes.c_void_p(dbg.data_ptr()))
        ws = eng.workspace(B, True)
        for _ in range(3):
            lib.vame_debug_gru_sweep(ctypes.byref(eng.dims), B, 0, L.ptr(eng.flat), L.ptr(eng.packed), L.ptr(ws), ws.numel(), L.cur_stream())
        torch.cuda.synchronize()
        lib.vame_set_debug_buffer(None)
        st = dbg.cpu().tolist()
        names = ["start", "prologue done", "after pdl_wait", "first chunk landed", "mma issued", "mma done", "epilogue+stores", "end",
                 "tmem loaded", "math done", "h stores issued"]
        print("[%s pdl=%d] fwd step kernel timeline (ns since start of last launch): " % (mode, pdl) +
              ", ".join("%s=%d" % (n, st[i] - st[0]) for i, n in enumerate(names)), flush=True)
        if lib.vame_get_option(b"rw") & 1:
            # rw forward kernel, step 10 of CTA 0: 0 loop top, 1 after cluster wait, 2 MMAs issued (MMA lane), 3 last gate's MMAs done,
            # 4 gate math done, 5 h pushed to the peers, 6 after cluster arrive, 7 global stores + next gi loads issued
            print("[%s pdl=%d] rw fwd stamps (ns since loop top): %s" % (mode, pdl, [st[i] - st[0] for i in range(12)]), flush=True)
        dbg.zero_()
        lib.vame_set_debug_buffer(ctypes.c_void_p(dbg.data_ptr()))
        lib.vame_debug_gru_sweep(ctypes.byref(eng.dims), B, 1, L.ptr(eng.flat), L.ptr(eng.packed), L.ptr(ws), ws.numel(), L.cur_stream())
        torch.cuda.synchronize()
        lib.vame_set_debug_buffer(None)
        st = dbg.cpu().tolist()
        names = ["loop top", "after cluster wait", "parts summed + math + smem A written", "after syncthreads", "mma issued",
                 "before mma wait", "mma done", "parts stored", "after cluster arrive", "gate-grad stores issued"]
        print("[%s pdl=%d] bwd persistent kernel, step 10 timeline (ns): " % (mode, pdl) +
              ", ".join("%s=%d" % (n, st[i] - st[0]) for i, n in enumerate(names)), flush=True)
        if lib.vame_get_option(b"rw") & 2:
            # rw backward kernel, step 10 of CTA 0: 0 loop top, 1 after cluster wait, 2 gate gradients + operand written,
            # 3 MMAs issued (MMA lane), 4 partial sums pushed, 5 after cluster arrive, 6 next step's loads issued
            print("[%s pdl=%d] rw bwd stamps (ns since loop top): %s" % (mode, pdl, [st[i] - st[0] for i in range(7)]), flush=True)
        print("[%s pdl=%d] rw=%d rw2=%d rw_timeouts=%d info=%s" % (mode, pdl, lib.vame_get_option(b"rw"), lib.vame_get_option(b"rw2"), lib.vame_get_option(b"rw_timeouts"),
              [lib.vame_get_option(b"rw_timeout_info%d" % i) for i in range(6)]), flush=True)
        e0.record()
        for _ in range(20):
            lib.vame_debug_gru_sweep(ctypes.byref(eng.dims), B, 0, L.ptr(eng.flat), L.ptr(eng.packed), L.ptr(ws), ws.numel(), L.cur_stream())
        e1.record()
        torch.cuda.synchronize()
        print("[%s pdl=%d] fwd sweep: %.2f us per step kernel" % (mode, pdl, e0.elapsed_time(e1) * 1e3 / (20 * T)), flush=True)
        e0.record()
        for _ in range(20):
            lib.vame_debug_gru_sweep(ctypes.byref(eng.dims), B, 1, L.ptr(eng.flat), L.ptr(eng.packed), L.ptr(ws), ws.numel(), L.cur_stream())
        e1.record()
        torch.cuda.synchronize()
        print("[%s pdl=%d] bwd sweep: %.2f us per step kernel" % (mode, pdl, e0.elapsed_time(e1) * 1e3 / (20 * T)), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child(sys.argv[2], int(sys.argv[3]), sys.argv[4], int(sys.argv[5]) if len(sys.argv) > 5 else 1,
              int(sys.argv[6]) if len(sys.argv) > 6 else 1, int(sys.argv[7]) if len(sys.argv) > 7 else 1)
        sys.exit(0)
    workload = sys.argv[1] if len(sys.argv) > 1 else "c2"
    for mode, pdl, streams, pers, flags in (("eager", 1, 1, 2, 0), ("graph", 1, 1, 2, 0)):
        print("##### variant mode=%s pdl=%d streams=%d persistent=%d flags=%d" % (mode, pdl, streams, pers, flags), flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "child", mode, str(pdl), workload, str(streams), str(pers), str(flags)], timeout=150,
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            print(r.stdout[-3000:], flush=True)
            print("variant %s pdl=%d rc=%d" % (mode, pdl, r.returncode), flush=True)
        except subprocess.TimeoutExpired as ex:
            out = ex.stdout.decode() if isinstance(ex.stdout, bytes) else (ex.stdout or "")
            print(out[-3000:], flush=True)
            print("variant %s pdl=%d TIMEOUT" % (mode, pdl), flush=True)
