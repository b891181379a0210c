#!/usr/bin/env python
"""bench.py — pose-windows/s of the RNN-VAE train step (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps 50 --warmup 10                      # our CUDA path (one JSON line)
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      # weak scaling, one rank per GPU, NCCL allreduce
    python bench.py --impl reference --steps 5 --warmup 1                 # the reference's CPU arithmetic (oracle port)

A "step" = forward + 4 losses + backward + [gradient allreduce] + AMSGrad over one batch of synthetic pose windows
(SURVEY.md §8d).  `value` is measured with the batch resident in HBM; `e2e` repeats it through host buffers (pinned
H2D copy of the batch and D2H read of the loss inside the timed region).  See DESIGN.md §Measurement.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (F, T, Z, H, future, S, per-GPU batch)
    "c2": (24, 30, 30, 256, False, 0, 256),      # BASELINE configs[1]: the configuration `metric` is quoted on
    "c2fut": (24, 30, 30, 256, True, 15, 256),
    "c3": (60, 60, 50, 256, True, 30, 512),      # BASELINE configs[2]
    "c5": (24, 30, 30, 256, False, 0, 512),      # BASELINE configs[4]: 512 windows per GPU (global 4096 at 8 GPUs)
    # BASELINE configs[3]: embedd_latent_vectors over a 1e6-frame synthetic series (inference; "batch" = frames of the series)
    "c4": (24, 30, 30, 256, True, 15, 1_000_000),
}
EMBED_FLOPS_PER_WINDOW = 96.71e6                 # encoder 96.58 M + both Lambda heads 0.12 M


def fwd_flops_per_window(F, T, Z, H, fut, S):
    """Algorithmic forward FLOPs per window, MAC = 2 FLOP, reference-as-written."""
    G = 3 * H
    enc = 2 * T * (F * G + H * G) * 2 + 2 * T * (2 * H * G + H * G) * 2
    lam = 2 * (4 * H) * Z * 2
    dec = Z * 2 * H * 2 + 2 * T * (Z * G + H * G) * 2 + T * 2 * H * F * 2
    f = enc + lam + dec
    if fut:
        f += Z * 2 * H * 2 + 2 * S * (Z * G + H * G) * 2 + S * 2 * H * F * 2
    return f


class ClockSampler:
    """nvidia-smi clocks/throttle sampling DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "10"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def wait_first(self, timeout=3.0):
        """nvidia-smi needs a moment to come up: block until its first sample so that short timed regions are covered."""
        t0 = time.time()
        while self.proc is not None and not self.lines and time.time() - t0 < timeout:
            time.sleep(0.01)

    def stop(self, t0=None, t1=None):
        """t0/t1 (time.time()): keep the samples that arrived inside the timed region; the sampler is started before the warm-up
        steps (nvidia-smi needs ~0.1 s to come up, longer than a short timed region), so if none fell inside, the samples taken
        under the same load during warm-up are used and `window` says so."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        window = "all"
        lines = [ln for _, ln in self.lines]
        if t0 is not None and t1 is not None:
            # the timed region plus the 0.15 s of identical load right before it (see main(): 150 extra warm-up steps)
            inside = [ln for ts, ln in self.lines if t0 - 0.15 <= ts <= t1 + 0.05]
            window = "timed region + the 0.15 s of identical load before it" if inside else "warm-up + timed region (same load)"
            lines = inside or lines
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2])); pw.append(float(p[3]))
            except ValueError:
                continue
            for n, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "window": window, "reasons": sorted(reasons)}


def usable_cpus():
    """Host threads this process can really use: scheduler affinity, capped by the cgroup CPU quota (if any) and by 64
    (ATen's GRU/GEMM sizes here stop scaling long before that; oversubscribing a quota-limited container is far slower)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q = open("/sys/fs/cgroup/cpu.max").read().split()
        if q[0] != "max":
            n = min(n, max(1, int(float(q[0]) / float(q[1]))))
    except Exception:
        pass
    return max(1, min(n, 64))


def cpu_port_step_time(F, T, Z, H, fut, S, B, steps, warmup, threads=None, budget_s=60.0):
    """The reference's CPU arithmetic (oracle torch port: same ATen calls) for the same step on the host cores.
    Bounded: stops early once `budget_s` seconds of CPU work have been spent."""
    import torch
    from oracle import vame_oracle as vo
    threads = threads or usable_cpus()
    torch.set_num_threads(threads)
    torch.manual_seed(19)
    port = vo.RefPort(2 * T, Z, F, fut, S, hidden=H)
    opt = vo.make_optimizer(port)
    x, xf, eps = vo.synthetic_batch(B, T, F, max(S, 1), Z)
    hp = dict(beta=1.0, kl_weight=1.0, kmeans_loss=Z, kmeans_lambda=0.1, bsize=B)
    t_start = time.perf_counter()
    for _ in range(warmup):
        vo.train_step(port, x, xf[:, :S] if fut else xf, eps, hp, optimizer=opt)
        if time.perf_counter() - t_start > budget_s / 2:
            break
    t0 = time.perf_counter()
    done = 0
    for _ in range(max(steps, 1)):
        vo.train_step(port, x, xf[:, :S] if fut else xf, eps, hp, optimizer=opt)
        done += 1
        if time.perf_counter() - t_start > budget_s:
            break
    dt = (time.perf_counter() - t0) / done
    return dt, torch.get_num_threads()


def train_epoch_throughput(F, T, Z, H, fut, S, B, dev, n_batches=200):
    """windows/s of the drop-in vame_b200.rnn_vae.train() (the function vame.train_model calls once per epoch) over n_batches
    batches, fed (a) by host float64 batches like the reference's DataLoader yields and (b) by the device sampler."""
    import tempfile
    import numpy as np
    import torch
    from vame_b200 import rnn_vae as rv
    from vame_b200.dataloader import SEQUENCE_DATASET, Data
    from vame_b200.rnn_model import RNN_VAE
    torch.manual_seed(19)
    model = RNN_VAE(2 * T, Z, F, fut, S, H, H, H, H, 0, 0, 0, False).cuda(dev)
    opt = torch.optim.Adam(model.parameters(), lr=5e-4, amsgrad=True)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=100, gamma=1)
    rng = np.random.default_rng(19)
    N = n_batches * B
    series = np.cumsum(rng.standard_normal((F, N)), axis=1) * 0.05 + rng.standard_normal((F, N))
    out = {"batches": n_batches, "unit": "windows/s"}
    with tempfile.TemporaryDirectory() as tmp:
        d = tmp + os.sep
        import contextlib
        import io
        np.save(d + "train_seq.npy", series)
        with contextlib.redirect_stdout(io.StringIO()):          # (the dataset prints like the reference's)
            ds = SEQUENCE_DATASET(d, data="train_seq.npy", train=True, temporal_window=2 * T)
        loader = Data.DataLoader(ds, batch_size=B, shuffle=True, drop_last=True)            # -> DeviceWindowSampler
        g = torch.Generator().manual_seed(3)
        host = [torch.randn(B, F, 2 * T, generator=g, dtype=torch.float64) for _ in range(8)]
        host_loader = [host[i % 8] for i in range(n_batches)]
        for name, ld in (("device_sampler", loader), ("host_f64_loader", host_loader)):
            with contextlib.redirect_stdout(io.StringIO()):
                rv.train(ld if name == "device_sampler" else host_loader[:8], 1, model, opt, "linear", 1, 0, 4, 2 * T, fut, S, sched,
                         "sum", "sum", Z, 0.1, B, False)                                # warm-up: graph capture
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                rv.train(ld, 2, model, opt, "linear", 1, 0, 4, 2 * T, fut, S, sched, "sum", "sum", Z, 0.1, B, False)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
            out[name] = len(ld) * B / dt
            out[name + "_ms_per_batch"] = dt / len(ld) * 1e3
    return out


def cudnn_reference_step(F, T, Z, H, fut, S, B, dev, steps=10, warmup=3):
    """The comparator BASELINE.md section 3 names: the reference's own modules (oracle port = the same nn.GRU / nn.Linear /
    torch.svd calls) moved to the B200 with .cuda(), i.e. PyTorch's cuDNN RNN path - the only other sm_100 kernel for this op."""
    import torch
    from oracle import vame_oracle as vo
    torch.manual_seed(19)
    port = vo.RefPort(2 * T, Z, F, fut, S, hidden=H)
    device = torch.device("cuda", dev)
    for m in port.mods.values():
        m.to(device)
    opt = vo.make_optimizer(port)
    x, xf, eps = (t.to(device) for t in vo.synthetic_batch(B, T, F, max(S, 1), Z))
    xf = xf[:, :S] if fut else xf
    hp = dict(beta=1.0, kl_weight=1.0, kmeans_loss=Z, kmeans_lambda=0.1, bsize=B)
    try:
        for _ in range(warmup):
            vo.train_step(port, x, xf, eps, hp, optimizer=opt)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            vo.train_step(port, x, xf, eps, hp, optimizer=opt)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"value": B / (ms * 1e-3), "unit": "windows/s", "ms_per_step": ms, "steps": steps,
                "what": "reference modules on the GPU (torch %s: cuDNN GRU, cuBLAS linears, cuSOLVER B x B SVD of cluster_loss, "
                        "torch.optim.Adam(amsgrad)), eager, fp32, .item() of 4-5 loss terms per step as in train()" % torch.__version__}
    except Exception as ex:      # a comparator must never take the benchmark down
        return {"value": None, "error": repr(ex)[:200]}


def bench_embed(args, F, T, Z, H, fut, S, n_frames):
    """BASELINE configs[3]: the sliding-window embedding (vame/analysis/pose_segmentation.py:87-98) of one n_frames-long series.
    A step = one pass over the whole series (N - T windows).  value: series and latent vectors resident in HBM (Engine.embed);
    e2e: the plugin's own function, vame_b200.pose_segmentation.embed_series - host (F, N) float64 array in, host (N - T, Z)
    float32 array out, copies inside the timed region.  N > 1: the window range is sharded over the ranks, no collective."""
    import numpy as np
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    K, W = max(args.steps if args.steps != 50 else 3, 1), max(args.warmup if args.warmup != 10 else 3, 3)
    n_win = n_frames - T
    cfg_desc = {"workload": "c4: embedd_latent_vectors over a %d-frame synthetic series, F=%d T=%d Z=%d H=%d (%d stride-1 windows, "
                            "encoder + Lambda mean), window range sharded over the GPUs" % (n_frames, F, T, Z, H, n_win),
                "windows": n_win, "parallelism": "dp%d (no collective on the data path)" % world,
                "l2": "per-pass working set (per-frame input projections 6.1 GB + per-chunk activations) exceeds the 126 MB L2"}
    if args.impl == "reference":
        if rank != 0:
            return 0
        import torch
        from oracle import vame_oracle as vo
        torch.manual_seed(19)
        port = vo.RefPort(2 * T, Z, F, fut, S, hidden=H)
        thr = usable_cpus()
        torch.set_num_threads(thr)
        rng = np.random.default_rng(5)
        n_s = 3000                                            # bounded sample: the literal batch-1 loop on the first n_s windows
        series = rng.standard_normal((F, n_s + T))
        for _ in range(max(W, 1) - 1):
            vo.embed_loop(port, series[:, :300 + T], T)
        t0 = time.perf_counter()
        for _ in range(K):
            vo.embed_loop(port, series, T)
        dt = (time.perf_counter() - t0) / K
        v = n_s / dt
        line = {"impl": "reference", "metric": "pose_windows_per_sec_embedd_latent_vectors", "value": v, "unit": "windows/s",
                "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": dict(cfg_desc, parallelism="cpu"),
                "cpu_baseline": {"value": v, "unit": "windows/s", "cores": thr, "kind": "port",
                                 "sample": "the reference's literal batch-1 loop (oracle port: same ATen calls) over the first %d "
                                           "windows of the series, %d passes; extrapolates linearly to the %d windows" % (n_s, K, n_win)},
                "e2e": {"value": v, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the vame_b200 hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from oracle import vame_oracle as vo     # only for the reference initial weights
    from vame_b200 import _lib as L
    from vame_b200 import pose_segmentation as ps
    from vame_b200.rnn_model import RNN_VAE
    for kv in args.opt:
        k_, v_ = kv.split("=")
        L.check(L.lib().vame_set_option(k_.encode(), int(v_)), "vame_set_option")
    if args.opt:
        cfg_desc["options"] = list(args.opt)
    torch.manual_seed(19)
    model = RNN_VAE(2 * T, Z, F, fut, S, H, H, H, H, 0, 0, 0, False).cuda(local_rank).eval()
    eng = model.engine
    rng = np.random.default_rng(5)
    series = rng.standard_normal((F, n_frames))               # (F, N) float64 like <file>-PE-seq-clean.npy
    per = (n_win + world - 1) // world
    first = min(n_win, rank * per)
    count = min(per, n_win - first)
    dev = torch.from_numpy(np.ascontiguousarray(series.T)).float().cuda()
    out = torch.empty(count, Z, device="cuda")
    lib = L.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.wait_first()
    for _ in range(W):
        eng.embed(dev, first_window=first, n_windows=count, out=out)
    barrier()
    l0 = lib.vame_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for _ in range(K):
        eng.embed(dev, first_window=first, n_windows=count, out=out)
    e1.record()
    barrier()
    t1 = time.time()
    clocks = sampler.stop(t0, t1)
    launches = lib.vame_launch_count() - l0
    ms = e0.elapsed_time(e1) / K
    # end to end: the plugin's function, host float64 (F, N) in -> host float32 (N - T, Z) out
    ps.embed_series(model, series, T, shard=(rank, world))
    barrier()
    tt0 = time.perf_counter()
    for _ in range(K):
        lat = ps.embed_series(model, series, T, shard=(rank, world))
    barrier()
    ms_e2e = (time.perf_counter() - tt0) / K * 1e3
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0].item()), float(t[1].item())
    if rank == 0:
        peak_burst, peak_sus, hbm, how = peaks()
        ach = EMBED_FLOPS_PER_WINDOW * n_win / (ms * 1e-3) / 1e12
        # parity spot check inside the benchmark: 512 random windows of this rank vs the CPU oracle port
        idx = np.sort(rng.choice(count, size=min(512, count), replace=False))
        port = vo.RefPort(2 * T, Z, F, fut, S, hidden=H).load_state_dict({k: v.cpu() for k, v in model.state_dict().items()})
        xw = torch.from_numpy(np.stack([series[:, first + i:first + i + T].T for i in idx])).float()
        with torch.no_grad():
            ref = port.lmbda(port.encode(xw), None)[1]
        err = float((torch.from_numpy(lat[idx]) - ref).abs().max() / ref.abs().max())
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            thr = usable_cpus()
            torch.set_num_threads(thr)
            n_s = 2000
            tc = time.perf_counter()
            vo.embed_loop(port, series[:, :n_s + T], T)
            dt = time.perf_counter() - tc
            cpu = {"value": n_s / dt, "unit": "windows/s", "cores": thr, "kind": "port",
                   "sample": "the reference's literal batch-1 loop over the first %d windows (pose_segmentation.py:87-98)" % n_s}
        line = {"metric": "pose_windows_per_sec_embedd_latent_vectors", "value": n_win / (ms * 1e-3), "unit": "windows/s",
                "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32 (tensor-core products as 3-pass bf16 hi/lo split, fp32 accumulate)", "data": "synthetic",
                "config": cfg_desc, "clocks": clocks,
                "e2e": {"value": n_win / (ms_e2e * 1e-3), "unit": "windows/s", "ms_per_step": ms_e2e,
                        "h2d_bytes_per_step": int(n_frames) * F * 4, "d2h_bytes_per_step": int(n_win) * Z * 4,
                        "what": "vame_b200.pose_segmentation.embed_series: host (F, N) float64 array -> host (N - T, Z) float32 array"},
                "gpu_launches": int(launches),
                "roofline": {"kernel": "whole embedding pass (recurrent sweeps + layer-1 input-projection GEMM)", "bound": "tensor",
                             "achieved": ach, "peak": peak_sus, "unit": "TFLOP/s", "frac": ach / peak_sus, "traffic": None,
                             "peak_source": how + ", sustained figure (seconds-long pass)",
                             "algorithmic_flops_per_window": EMBED_FLOPS_PER_WINDOW, "issued_frac": 3 * ach / peak_sus,
                             "note": "algorithmic FLOPs as the reference computes them (SURVEY 8d); the kernels issue 3x that in bf16 "
                                     "MMAs (hi/lo split), so the algorithmic ceiling is 1/3 of the bf16 peak"},
                "cpu_baseline": cpu, "max_rel_err_vs_oracle_512_windows": err}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 1590.0, 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--per-gpu-batch", type=int, default=0)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train-epoch", action="store_true", help="skip the train()-epoch measurement (e2e_train_epoch)")
    ap.add_argument("--no-cudnn", action="store_true", help="skip the cuDNN comparator (the reference's modules moved to the GPU)")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE",
                    help="library option for experiments (vame_set_option), recorded in config.options")
    args = ap.parse_args()
    F, T, Z, H, fut, S, B = WORKLOADS[args.workload]
    if args.per_gpu_batch:
        B = args.per_gpu_batch
    if args.workload == "c4":
        return bench_embed(args, F, T, Z, H, fut, S, B)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    W = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    K = args.steps
    cfg_desc = {"workload": "%s: synthetic pose windows F=%d T=%d Z=%d H=%d, %d windows/GPU, %s, train step = fwd + rec/KL/k-means%s losses + "
                            "bwd + %sAMSGrad" % (args.workload, F, T, Z, H, B, "future decoder S=%d" % S if fut else "no future decoder",
                                                 "/future" if fut else "", "NCCL grad allreduce + " if world > 1 else ""),
                "global_batch": B * world, "per_gpu_batch": B, "parallelism": "dp%d" % world,
                "l2": "per-step working set (activations + P16 operand copies, > 1 GB) exceeds the 126 MB L2; no explicit flush",
                "batches": "8 distinct synthetic batches in rotation (device-resident for `value`, pinned host memory for `e2e`)"}
    step_flops = 3 * fwd_flops_per_window(F, T, Z, H, fut, S) * B            # train = 3 x forward

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return 0
        dt, thr = cpu_port_step_time(F, T, Z, H, fut, S, B, K, W)
        v = B / dt
        line = {"impl": "reference", "metric": "pose_windows_per_sec_train_step", "value": v, "unit": "windows/s", "n_gpus": args.gpus,
                "steps": K, "warmup": W, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": dict(cfg_desc, parallelism="cpu"),
                "cpu_baseline": {"value": v, "unit": "windows/s", "cores": thr, "kind": "port",
                                 "sample": "%d steps of the same %d-window batch on the host CPU (oracle torch port = the reference's ATen calls)" % (K, B)},
                "e2e": {"value": v, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the vame_b200 hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from oracle import vame_oracle as vo     # only for the deterministic synthetic inputs / reference init
    from vame_b200.engine import Engine, TrainStep
    from vame_b200 import _lib as L
    import ctypes

    for kv in args.opt:
        k_, v_ = kv.split("=")
        L.check(L.lib().vame_set_option(k_.encode(), int(v_)), "vame_set_option")
    if args.opt:
        cfg_desc["options"] = list(args.opt)
    torch.manual_seed(19)
    port = vo.RefPort(2 * T, Z, F, fut, S, hidden=H)                  # reference default init under seed 19 (rnn_vae.py:292)
    eng = Engine(F, T, Z, H, H, H, fut, S, False, device="cuda:%d" % local_rank)
    eng.load_state_dict(port.state_dict())
    # NB distinct synthetic batches in rotation (a fresh batch every step, like training): with ONE repeated batch the latent
    # Gram matrix barely moves between steps and the warm-started eigen-solver of the k-means prior would look better than in use
    NB = 8
    batches = []
    for i in range(NB):
        bx, bf, be = vo.synthetic_batch(B, T, F, max(S, 1), Z, seed=19 + rank + 1000 * i)
        batches.append((bx, bf[:, :S] if fut else None, be))
    x, xf, eps = batches[0]
    dev_batches = [(a.cuda(), b.cuda() if fut else None, c.cuda()) for a, b, c in batches]
    cfg = eng.loss_cfg(kmeans_loss=Z, kmeans_lambda=0.1, bsize=B, beta=1.0, kl_weight=1.0)
    eng.set_hyper(lr=5e-4, kl_weight=1.0, beta=1.0, kmeans_lambda=0.1)
    ts = TrainStep(eng, B, cfg, world=world, use_graph=not args.no_graph)
    graphed = ts.capture()
    ts.load(x.cuda(), xf.cuda() if fut else None, eps.cuda())
    lib = L.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.wait_first()                                   # the first sample is in before the warm-up starts
    def step(i):
        ts.load(*dev_batches[i % NB])                      # device -> static buffers (three small copies on the stream)
        return ts.run()

    for i in range(W):
        step(i)
    barrier()
    # keep the GPU under the same load for a few sampling periods (10 ms each) before the timed region, so that the clocks line
    # always holds samples taken under load even when the timed region itself is shorter than one period.  A FIXED number of
    # steps: with N > 1 every step contains collectives, so all ranks must run the same count.
    for i in range(150):
        step(i)
    barrier()
    launches0 = lib.vame_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    for i in range(K):
        step(i)
    e1.record()
    barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1)
    ms = e0.elapsed_time(e1) / K
    launches_eager = lib.vame_launch_count() - launches0
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    loss_after = float(ts.losses[4].item())

    # ---- end to end through host buffers: pinned H2D of the batch + D2H of the loss every step
    host_batches = [(a.pin_memory(), b.pin_memory() if fut else None, c.pin_memory()) for a, b, c in batches]
    xh, fh, eh = host_batches[0]
    loss_h = torch.zeros(8).pin_memory()
    # the host batch of step i+1 is handed to TrainStep.load() right after step i has been enqueued: it is uploaded on a copy
    # stream while step i computes; every timed step still contains one H2D of a full batch and one D2H of the loss vector
    for i in range(3):
        ts.load(*host_batches[i % NB])
        ts.run()
    barrier()
    e0.record()
    ts.load(*host_batches[0])
    for i in range(K):
        out = ts.run()
        if i + 1 < K:
            ts.load(*host_batches[(i + 1) % NB])
        loss_h.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()           # the caller consumes the loss every step
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1) / K
    if world > 1:
        t = torch.tensor([ms_e2e], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    h2d = xh.numel() * 4 + eh.numel() * 4 + (fh.numel() * 4 if fut else 0)

    # ---- launches per step: count one eager step (graph replays do not pass through the host launchers)
    if graphed:
        c0 = lib.vame_launch_count()
        ts._phase1(); ts._phase2()
        torch.cuda.synchronize()
        per_step = lib.vame_launch_count() - c0
        gpu_launches = per_step * K
    else:
        gpu_launches = launches_eager

    # ---- roofline of the dominant kernels: the recurrent step kernels, timed alone with CUDA events on their stream
    roof = None
    if rank == 0:
        ts._phase1()                                       # leaves forward + backward state in the workspace
        ws = eng.workspace(B, True)
        res = {}
        pers = lib.vame_get_option(b"persistent")
        rw, rw2 = lib.vame_get_option(b"rw"), lib.vame_get_option(b"rw2")
        tiles = (B + 127) // 128
        # same rule as rw_applicable() / rw_groups_per_cluster() in csrc/gru_rw.cu: 16-row clusters while they fit one wave,
        # otherwise (H = 256) two interleaved 16-row groups per cluster
        ng = lib.vame_get_option(b"rw_ng")
        ng = ng if ng in (1, 2) else (1 if tiles * 64 <= 132 else 2)
        ng = ng if (H == 256 and rw2) else 1
        rw_ok = H % 64 == 0 and 64 <= H <= 256 and tiles * 64 // ng <= 132 * lib.vame_get_option(b"rw_waves")
        rw_name = "gru_rw2_%s_kernel (per time step)" if (H == 256 and rw2) else "gru_rw_%s_kernel (per time step)"
        names = (rw_name % "fwd" if (rw & 1) and rw_ok else "gru_seq_fwd_kernel (per time step)" if pers & 1 else "gru_step_fwd_kernel",
                 rw_name % "bwd" if (rw & 2) and rw_ok else "gru_seq_bwd_kernel (per time step)" if pers & 2 else "gru_step_bwd_kernel")
        for which, name in ((0, names[0]), (1, names[1])):
            for _ in range(3):
                L.check(lib.vame_debug_gru_sweep(ctypes.byref(eng.dims), B, which, L.ptr(eng.flat), L.ptr(eng.packed), L.ptr(ws), ws.numel(), L.cur_stream()), "sweep")
            torch.cuda.synchronize()
            reps = 20
            e0.record()
            for _ in range(reps):
                L.check(lib.vame_debug_gru_sweep(ctypes.byref(eng.dims), B, which, L.ptr(eng.flat), L.ptr(eng.packed), L.ptr(ws), ws.numel(), L.cur_stream()), "sweep")
            e1.record()
            torch.cuda.synchronize()
            res[name] = e0.elapsed_time(e1) * 1e-3 / (reps * T)      # seconds per launch (both directions in one launch)
        peak_burst, peak_sus, hbm, how = peaks()
        flops_launch = 2 * B * (3 * H) * H * 2                      # both directions: [B,H] x [H,3H] MACs, MAC = 2 FLOP
        dom = max(res, key=res.get)
        ach = flops_launch / res[dom] / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")      # dram bytes per launch from the committed ncu --set full capture
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            traffic = tj.get(dom.split(" ")[0] + ("@2groups" if ng == 2 else ""), tj.get(dom.split(" ")[0]))
        roof = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": peak_burst, "unit": "TFLOP/s", "frac": ach / peak_burst,
                "traffic": traffic, "peak_source": how + ", burst figure (kernel timed alone)",
                "algorithmic_flops_per_launch": flops_launch, "issued_mma_flops_per_launch": 4 * flops_launch,
                "issued_frac": 4 * ach / peak_burst, "algorithmic_ceiling_frac": 0.25,
                "us_per_launch": {k: v * 1e6 for k, v in res.items()}, "row_groups_per_cluster": ng,
                "note": "algorithmic FLOPs = one fp32 recurrent projection per direction per time step; the sweep kernels issue 4x that "
                        "in bf16 MMAs (hi/lo split of both operands in ONE MMA), so their algorithmic ceiling is 1/4 of the bf16 peak "
                        "(frac <= 0.25; issued_frac = 4 x frac); the GEMMs and the embedding kernel issue 3x (ceiling 1/3).  The step is "
                        "a serial chain of %d dependent time steps: per step and SM 48 MMAs at the 46-cycle issue floor (1.1 us), "
                        "~0.5 us for 12 KB of incoming DSMEM (~15 B/cycle/SM measured), ~0.4 us of tensor-memory drain (32 B/cycle per "
                        "lane quarter measured) and the gate math; with 512 windows per GPU two independent 16-row groups per cluster "
                        "take turns on the tensor pipe (DESIGN.md section 4.1, profiles/r2_tmem_dsmem_microbench.log)" % (6 * T),
                "step_tflops_algorithmic": step_flops / (ms * 1e-3) / 1e12,
                "step_frac_of_sustained_peak": step_flops / (ms * 1e-3) / 1e12 / peak_sus}

    # ---- the plugin call itself: windows/s of vame_b200.rnn_vae.train() over one epoch of >= 200 batches (rank 0, N = 1):
    #      (a) fed by a reference-style loader of HOST float64 (B, F, 2T) batches (rnn_vae.py:106-112: cast + H2D per batch),
    #      (b) fed by the on-device window sampler that install() binds into vame.model.rnn_vae (one graph replay per batch)
    epoch = None
    if rank == 0 and world == 1 and not args.no_train_epoch:
        epoch = train_epoch_throughput(F, T, Z, H, fut, S, B, local_rank)
    cudnn = None
    if rank == 0 and world == 1 and not args.no_cudnn:
        cudnn = cudnn_reference_step(F, T, Z, H, fut, S, B, local_rank)

    # ---- CPU baseline beside it (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        dt, thr = cpu_port_step_time(F, T, Z, H, fut, S, B, 3, 1)
        cpu = {"value": B / dt, "unit": "windows/s", "cores": thr, "kind": "port",
               "sample": "3 timed steps (1 warm-up) of the same %d-window train step on the host CPU, %d torch threads" % (B, thr),
               "ms_per_step": dt * 1e3}

    if rank == 0:
        line = {"metric": "pose_windows_per_sec_train_step", "value": B * world / (ms * 1e-3), "unit": "windows/s", "n_gpus": world,
                "steps": K, "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 (tensor-core products as 3-pass bf16 hi/lo split, fp32 accumulate)", "data": "synthetic",
                "config": dict(cfg_desc, cuda_graph=bool(graphed)), "clocks": clocks,
                "e2e": {"value": B * world / (ms_e2e * 1e-3), "unit": "windows/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": 32},
                "gpu_launches": int(gpu_launches), "roofline": roof, "cpu_baseline": cpu, "loss_after": loss_after,
                "e2e_train_epoch": epoch, "cudnn_reference": cudnn, "nccl_in_graph": bool(getattr(ts, "nccl_in_graph", False))}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
