"""world_size-2 `gloo` tests (CPU) of the data-parallel host logic:
  * the gradient exchange is ONE sum-allreduce over the flat buffer and the 1/world average is applied by the optimizer's
    grad_scale -> identical to "reference on each shard, DDP-style gradient average, one AMSGrad step" 
  * embedding shards the window range with no collective on the data path and the shards tile the range exactly."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import gru_numpy as gnp
from oracle import vame_oracle as vo


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _FakeEngine:
    """Stands in for vame_b200.engine.Engine on a CPU box: only the flat buffers matter to the exchange / the safeguards."""

    def __init__(self, grad, flat=None):
        self.grad = grad
        self.flat = flat
        self.opt_state = None
        self.dirty = False

    def init_optimizer(self):
        if self.opt_state is None:
            z = lambda: torch.zeros_like(self.flat)  # noqa: E731
            self.opt_state = dict(exp_avg=z(), exp_avg_sq=z(), max_exp_avg_sq=z(), step=torch.zeros(1, dtype=torch.int32))
        return self.opt_state

    def mark_dirty(self):
        self.dirty = True


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vame_b200.rnn_vae import allreduce_gradients, _world
    assert _world() == world
    B, T, F, Z, H, S = 8, 6, 5, 4, 32, 3
    torch.manual_seed(19)
    portm = vo.RefPort(2 * T, Z, F, True, S, hidden=H)               # identical weights on every rank
    x, xf, eps = vo.synthetic_batch(B, T, F, S, Z, seed=19 + rank)    # rank-specific shard (bench.py convention)
    hp = dict(beta=1.0, kl_weight=1.0, kmeans_loss=Z, kmeans_lambda=0.1, bsize=B)
    _, grads, _ = vo.train_step(portm, x, xf, eps, hp)
    names = [k for k, _ in portm.named_parameters()]
    flat = torch.cat([grads[k].reshape(-1) for k in names])
    eng = _FakeEngine(flat.clone())
    allreduce_gradients(eng)                                          # the product's exchange step
    # one AMSGrad step with grad_scale = 1/world (numpy restatement of the fused optimizer kernel's arithmetic)
    w = torch.cat([p.detach().reshape(-1) for _, p in portm.named_parameters()]).numpy().astype(np.float64)
    g = eng.grad.numpy().astype(np.float64) / world
    m, v, vm = np.zeros_like(w), np.zeros_like(w), np.zeros_like(w)
    gnp.amsgrad_step(w, g, m, v, vm, 1, 5e-4)
    # replica-consistency safeguards: ranks that start from different weights / optimizer state are detected, and
    # made identical by the broadcast that precedes the first data-parallel train step
    from vame_b200.rnn_vae import _broadcast_replicas, assert_replicas_consistent
    from vame_b200._lib import VameB200Error
    e2 = _FakeEngine(None, flat=torch.full((1000,), float(rank + 1)))
    e2.init_optimizer()["exp_avg"].fill_(float(rank))
    try:
        assert_replicas_consistent(e2)
        diverged_detected = False
    except VameB200Error:
        diverged_detected = True
    _broadcast_replicas(e2)
    assert_replicas_consistent(e2)                                    # passes after the broadcast
    ok = diverged_detected and bool((e2.flat == 1.0).all()) and bool((e2.opt_state["exp_avg"] == 0.0).all()) and e2.dirty
    q.put((rank, flat.numpy(), eng.grad.numpy(), w, ok))
    dist.destroy_process_group()


def test_dp_allreduce_equals_shard_average():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    local = [r[1] for r in res]
    summed = local[0] + local[1]
    for r in res:
        np.testing.assert_allclose(r[2], summed, rtol=1e-6, atol=1e-7)       # every rank holds the global sum
    np.testing.assert_array_equal(res[0][3], res[1][3])                      # replicas stay bit-identical after the step
    assert res[0][4] and res[1][4], "replica-consistency safeguards (broadcast / checksum) failed"
    # single-process emulation: reference on each shard, average, one step
    B, T, F, Z, H, S = 8, 6, 5, 4, 32, 3
    torch.manual_seed(19)
    portm = vo.RefPort(2 * T, Z, F, True, S, hidden=H)
    opt = vo.make_optimizer(portm)
    hp = dict(beta=1.0, kl_weight=1.0, kmeans_loss=Z, kmeans_lambda=0.1, bsize=B)
    acc = None
    for rank in range(world):
        x, xf, eps = vo.synthetic_batch(B, T, F, S, Z, seed=19 + rank)
        _, grads, _ = vo.train_step(portm, x, xf, eps, hp)
        acc = grads if acc is None else {k: acc[k] + grads[k] for k in grads}
    for k, p in portm.named_parameters():
        p.grad = acc[k] / world
    opt.step()
    ref = torch.cat([p.detach().reshape(-1) for _, p in portm.named_parameters()]).numpy()
    assert np.abs(res[0][3] - ref).max() < 2e-6


def test_embed_shards_tile_the_window_range():
    """pose_segmentation.embed_series(shard=(rank, world)) partitioning arithmetic (no GPU needed)."""
    def shard(n_win, rank, world):
        per = (n_win + world - 1) // world
        first = min(n_win, rank * per)
        return first, min(per, n_win - first)
    for n_win in (0, 1, 7, 29967, 999970):
        for world in (1, 2, 3, 8):
            spans = [shard(n_win, r, world) for r in range(world)]
            assert sum(c for _, c in spans) == n_win
            pos = 0
            for f, c in spans:
                assert f == pos or c == 0
                pos += c
