"""2-rank (or N-rank) CUDA check of the data-parallel train step, launched by tests/test_gpu_dp.py with torchrun:
every rank runs engine.TrainStep steps on its own shard (NCCL gradient allreduce inside the step); afterwards
  * the replicas' parameters must be BIT-IDENTICAL across ranks,
  * and equal to "the oracle on each shard, gradients averaged, one AMSGrad step"  within the
    optimizer tolerance of tests/test_gpu_model.py.
Prints one JSON line on rank 0."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from oracle import vame_oracle as vo
    from vame_b200.engine import Engine, TrainStep
    from vame_b200.rnn_vae import _broadcast_replicas, assert_replicas_consistent, replica_checksum
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B, T, F, Z, H, S = 48, 8, 10, 8, 256, 4
    torch.manual_seed(19 + 7 * rank)                     # deliberately DIFFERENT initial weights per rank ...
    port = vo.RefPort(2 * T, Z, F, True, S, hidden=H)
    eng = Engine(F, T, Z, H, H, H, True, S, False, device="cuda:%d" % local)
    eng.load_state_dict(port.state_dict())
    _broadcast_replicas(eng)                              # ... made identical by the broadcast of rank 0's
    assert_replicas_consistent(eng)
    torch.manual_seed(19)
    port0 = vo.RefPort(2 * T, Z, F, True, S, hidden=H)    # = rank 0's weights (seed 19 + 0)
    x, xf, eps = vo.synthetic_batch(B, T, F, S, Z, seed=19 + rank)
    cfg = eng.loss_cfg(kmeans_loss=Z, kmeans_lambda=0.1, bsize=B, beta=1.0, kl_weight=1.0)
    eng.set_hyper(lr=5e-4, kl_weight=1.0, beta=1.0, kmeans_lambda=0.1)
    ts = TrainStep(eng, B, cfg, world=world)
    graphed = ts.capture()
    ts.load(x.cuda(), xf.cuda(), eps.cuda())
    steps = 3
    for _ in range(steps):
        ts.run()
    torch.cuda.synchronize()
    assert_replicas_consistent(eng)                       # checksum equality (tol 0) across ranks
    flat = eng.flat.clone()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    identical = all(torch.equal(gathered[0], g) for g in gathered)
    # oracle: every shard on the CPU, averaged gradients, AMSGrad (rank 0 only)
    out = None
    if rank == 0:
        opt = vo.make_optimizer(port0)
        hp = dict(beta=1.0, kl_weight=1.0, kmeans_loss=Z, kmeans_lambda=0.1, bsize=B)
        for _ in range(steps):
            acc = None
            for r in range(world):
                xr, xfr, er = vo.synthetic_batch(B, T, F, S, Z, seed=19 + r)
                _, grads, _ = vo.train_step(port0, xr, xfr, er, hp)
                acc = grads if acc is None else {k: acc[k] + grads[k] for k in grads}
            for k, p in port0.named_parameters():
                p.grad = acc[k] / world
            opt.step()
        v = eng.views()
        worst, frac = 0.0, 0.0
        for k, p in port0.named_parameters():
            err = (v[k].cpu() - p.detach()).abs()
            worst = max(worst, float(err.max()))
            frac = max(frac, float((err > 5e-6).float().mean()))
        out = {"world": world, "graphed": bool(graphed), "single_graph": bool(ts.graphs is not None and ts.graphs[1] is None),
               "replicas_bit_identical": bool(identical), "max_abs_err_vs_oracle": worst, "frac_above_5e-6": frac,
               "checksum": replica_checksum(eng).tolist(), "ok": bool(identical and worst <= steps * 5e-4 + 1e-6 and frac <= 2e-3)}
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
