"""Host-side mirror of the hot-path functions of vame/model/rnn_vae.py: the four loss functions (:35-60),
kl_annealing (:63-81), gaussian (:84-91) and the per-epoch train (:94-164) / test (:167-210) loops, with the same
signatures and return values (including the reference's division by ``idx`` = last batch index).

``train`` / ``test`` drive the fused CUDA step (forward + losses + backward [+ NCCL allreduce] + AMSGrad) of
``vame_b200.engine.Engine``; loss scalars are accumulated on the device and read back once per epoch.
"""
import numpy as np
import torch
import torch.distributed as dist

from ._lib import VameB200Error
from .rnn_model import RNN_VAE


# ---- loss functions (generic autograd path; the fused step computes the same terms inside the library) ------------
def reconstruction_loss(x, x_tilde, reduction):
    """nn.MSELoss(reduction)(x_tilde, x)  [rnn_vae.py:35-38]"""
    d = x_tilde - x
    s = (d * d).sum()
    return s if reduction == "sum" else s / d.numel()


def future_reconstruction_loss(x, x_tilde, reduction):
    """[rnn_vae.py:40-43]"""
    return reconstruction_loss(x, x_tilde, reduction)


class _ClusterLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, latent, kloss, lmbda, batch_size):
        from . import _lib as L
        import ctypes
        lat = latent.detach().float().contiguous()
        B, Z = lat.shape
        loss = torch.zeros(1, dtype=torch.float64, device=lat.device)
        dl = torch.empty_like(lat)
        L.check(L.lib().vame_cluster_loss(L.ptr(lat), B, Z, int(kloss), float(lmbda), float(batch_size), 1.0, L.ptr(loss), L.ptr(dl),
                                          L.cur_stream()), "vame_cluster_loss")
        ctx.save_for_backward(dl)
        return loss[0].float()

    @staticmethod
    def backward(ctx, g):
        (dl,) = ctx.saved_tensors
        return g * dl, None, None, None


def cluster_loss(H, kloss, lmbda, batch_size):
    """lmbda * sum(sqrt(svd(H.T @ H / batch_size)[:kloss]))  [rnn_vae.py:45-50]; H is latent.T (Z, B) at the call sites
    (:126,:137).  Evaluated through the Z x Z Gram matrix + Jacobi eigensolver kernel (vame_cluster_loss)."""
    if not H.is_cuda:
        raise VameB200Error("vame_b200.cluster_loss needs CUDA tensors (no CPU fallback)")
    return _ClusterLossFn.apply(H.T, kloss, lmbda, batch_size)


def kullback_leibler_loss(mu, logvar):
    """-0.5 * mean(1 + logvar - mu^2 - exp(logvar))  [rnn_vae.py:53-60]"""
    return -0.5 * torch.mean(1 + logvar - mu.pow(2) - logvar.exp())


def kl_annealing(epoch, kl_start, annealtime, function):
    """[rnn_vae.py:63-81]"""
    if epoch > kl_start:
        if function == "linear":
            return min(1, (epoch - kl_start) / (annealtime))
        if function == "sigmoid":
            return float(1 / (1 + np.exp(-0.9 * (epoch - annealtime))))
        raise NotImplementedError('currently only "linear" and "sigmoid" are implemented')
    return 0


def gaussian(ins, is_training, seq_len, std_n=0.8):
    """input noise x + N(0,1) * std_n * std_t(x)  [rnn_vae.py:84-91] (cfg['noise'], default False)"""
    if is_training:
        emp_std = (ins.std(1) * std_n).unsqueeze(1)
        return ins + torch.randn_like(ins) * emp_std
    return ins


# ---- data-parallel helper ------------------------------------------------------------------------------------------
def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def allreduce_gradients(engine):
    """The single gradient exchange of the data-parallel step: one NCCL sum-allreduce over the flat fp32 gradient buffer
    (the 1/world average is folded into the optimizer kernel's grad_scale)."""
    if _world() > 1:
        dist.all_reduce(engine.grad, op=dist.ReduceOp.SUM)


def _engine_of(model):
    if not isinstance(model, RNN_VAE):
        raise VameB200Error("vame_b200.train/test need a vame_b200.rnn_model.RNN_VAE (got %s)" % type(model).__name__)
    return model.engine


def _train_step(model, eng, batch, cfg, world, betas, adam_eps, sampler=None):
    """Cached engine.TrainStep (captured CUDA graphs) for this batch size / loss configuration [/ device sampler];
    hyper-parameters that change between epochs (lr, kl_weight, beta, lambda) are read from device memory, so the graphs stay
    valid."""
    from .engine import TrainStep
    key = (int(batch), cfg.mse_red_mean, cfg.mse_pred_mean, cfg.kmeans_loss, float(cfg.bsize), cfg.with_future, int(world),
           tuple(betas), float(adam_eps), id(sampler) if sampler is not None else None)
    cache = model.__dict__.setdefault("_b200_steps", {})
    ts = cache.get(key)
    if ts is None:
        import copy
        if world > 1:
            _broadcast_replicas(eng)
        ts = TrainStep(eng, batch, copy.copy(cfg), world=world, betas=betas, eps=adam_eps, sampler=sampler)
        ts.capture()
        cache[key] = ts
    return ts


def _broadcast_replicas(eng):
    """Data-parallel replicas must start identical: rank 0's parameters and optimizer state are broadcast once, when the first
    train step of a world > 1 job is built (different seeds or a pretrained_model loaded on one rank only would otherwise
    diverge silently - only gradients are summed)."""
    eng.init_optimizer()
    dist.broadcast(eng.flat, src=0)
    for k in ("exp_avg", "exp_avg_sq", "max_exp_avg_sq", "step"):
        dist.broadcast(eng.opt_state[k], src=0)
    eng.mark_dirty()


def replica_checksum(eng):
    """(sum, sum of squares) of the flat parameters in float64: equal on every rank of a consistent data-parallel job."""
    f = eng.flat.double()
    return torch.stack([f.sum(), (f * f).sum()])


def assert_replicas_consistent(eng, tol=0.0):
    """Raises if the ranks' parameters differ (cheap: two scalars per rank)."""
    if _world() <= 1:
        return
    c = replica_checksum(eng)
    lo, hi = c.clone(), c.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if bool(((hi - lo).abs() > tol * hi.abs().clamp_min(1e-30)).any()):
        raise VameB200Error("data-parallel replicas have diverged: parameter checksums differ across ranks (%s vs %s)"
                            % (lo.tolist(), hi.tolist()))


def _to_device(data_item, seq_len_half, future_steps, device):
    """rnn_vae.py:107-112: (B, F, 2T) float64 loader item -> data (B,T,F), fut (B,S,F) float32 on the GPU."""
    data_item = data_item.permute(0, 2, 1)
    data = data_item[:, :seq_len_half, :].to(dtype=torch.float32)
    fut = data_item[:, seq_len_half:seq_len_half + future_steps, :].to(dtype=torch.float32)
    return data.to(device, non_blocking=True), fut.to(device, non_blocking=True)


def train(train_loader, epoch, model, optimizer, anneal_function, BETA, kl_start,
          annealtime, seq_len, future_decoder, future_steps, scheduler, mse_red,
          mse_pred, kloss, klmbda, bsize, noise):
    """One epoch  [rnn_vae.py:94-164].  Returns (kl_weight, train_loss/idx, kl_weight*kmeans/idx, kl/idx, mse/idx, fut/idx)."""
    model.train()
    eng = _engine_of(model)
    seq_len_half = int(seq_len / 2)
    kl_weight = kl_annealing(epoch, kl_start, annealtime, anneal_function)
    lr = optimizer.param_groups[0]["lr"]
    betas = optimizer.param_groups[0].get("betas", (0.9, 0.999))
    adam_eps = optimizer.param_groups[0].get("eps", 1e-8)
    cfg = eng.loss_cfg(mse_red, mse_pred, kloss, klmbda, bsize, BETA, kl_weight, with_future=bool(future_decoder))
    acc = torch.zeros(8, dtype=torch.float64, device=eng.device)
    world = _world()
    idx = -1
    loss_last = None
    eng.set_hyper(lr=lr, kl_weight=kl_weight, beta=BETA, kmeans_lambda=klmbda)
    from .dataloader import DeviceWindowSampler
    fused_sampler = isinstance(train_loader, DeviceWindowSampler) and not (noise == True)   # noqa: E712
    if fused_sampler:
        # the sampler's launch is the first node of the captured step: one CUDA-graph replay per batch, no host work per window
        ts = _train_step(model, eng, train_loader.batch_size, cfg, world, betas, adam_eps, sampler=train_loader)
        ts.acc.zero_()
        n_batches = len(train_loader)
        for idx in range(n_batches):
            loss_last = ts.run()
        acc = ts.acc
    for idx, data_item in (enumerate(train_loader) if not fused_sampler else ()):
        data, fut = _to_device(data_item, seq_len_half, future_steps, eng.device)
        eps = torch.randn(data.shape[0], eng.dims.zdims, device=eng.device)
        if noise == True:   # noqa: E712
            # with input noise the model sees data_gaussian but the loss target stays the clean data (rnn_vae.py:116-124)
            eng.forward(gaussian(data, True, seq_len_half), eps, save=True, want=())
            losses = eng.loss(cfg, fut if future_decoder else None, want_grads=True, target=data)
            eng.backward(cfg)
            allreduce_gradients(eng)
            eng.adam_step(lr=lr, betas=betas, eps=adam_eps, grad_scale=1.0 / world)
        else:
            # steady state: the whole step is two CUDA-graph replays around the (optional) NCCL allreduce
            ts = _train_step(model, eng, data.shape[0], cfg, world, betas, adam_eps)
            ts.load(data, fut if future_decoder else None, eps)
            losses = ts.run()
        acc += losses.double()
        loss_last = losses
    if idx < 0:
        raise ValueError("empty train_loader")
    model.bind_flat_grads()
    a = acc.cpu().tolist()                       # the only device->host sync of the epoch
    rec, fl, kl, km, total = a[0], a[1], a[2], a[3], a[4]
    sched_loss = loss_last[4]                    # rnn_vae.py:155: the scheduler sees the LAST batch's loss
    if world > 1:
        # every rank must take the same lr decision (ReduceLROnPlateau): feed the rank-mean of the last-batch loss, and make sure
        # the replicas still agree (a silent divergence would otherwise only show up as a bad model)
        sched_loss = sched_loss.clone()
        dist.all_reduce(sched_loss, op=dist.ReduceOp.SUM)
        sched_loss /= world
        assert_replicas_consistent(eng)
    scheduler.step(sched_loss)
    div = idx if idx > 0 else float("nan")       # rnn_vae.py:157-164 divide by the last batch index (reference quirk)
    if future_decoder:
        print('Train loss: {:.3f}, MSE-Loss: {:.3f}, MSE-Future-Loss {:.3f}, KL-Loss: {:.3f}, Kmeans-Loss: {:.3f}, weight: {:.2f}'.format(
            total / div, rec / div, fl / div, BETA * kl_weight * kl / div, kl_weight * km / div, kl_weight))
    else:
        print('Train loss: {:.3f}, MSE-Loss: {:.3f}, KL-Loss: {:.3f}, Kmeans-Loss: {:.3f}, weight: {:.2f}'.format(
            total / div, rec / div, BETA * kl_weight * kl / div, kl_weight * km / div, kl_weight))
    return kl_weight, total / div, kl_weight * km / div, kl / div, rec / div, fl / div


def test(test_loader, epoch, model, optimizer, BETA, kl_weight, seq_len, mse_red, kloss, klmbda, future_decoder, bsize):
    """Evaluation epoch  [rnn_vae.py:167-210].  Returns (mse/idx, loss/idx, kl_weight*kmeans)."""
    model.eval()
    eng = _engine_of(model)
    seq_len_half = int(seq_len / 2)
    cfg = eng.loss_cfg(mse_red, "sum", kloss, klmbda, bsize, BETA, kl_weight, with_future=False)
    acc = torch.zeros(8, dtype=torch.float64, device=eng.device)
    idx = -1
    with torch.no_grad():
        for idx, data_item in enumerate(test_loader):
            data, _ = _to_device(data_item, seq_len_half, 0, eng.device)
            eng.forward(data, None, save=False, want=())
            acc += eng.loss(cfg, None, want_grads=False).double()
    if idx < 0:
        raise ValueError("empty test_loader")
    a = acc.cpu().tolist()
    rec, kl, km, total = a[0], a[2], a[3], a[4]
    div = idx if idx > 0 else float("nan")
    print('Test loss: {:.3f}, MSE-Loss: {:.3f}, KL-Loss: {:.3f}, Kmeans-Loss: {:.3f}'.format(
        total / div, rec / div, BETA * kl_weight * kl / div, kl_weight * km / div))
    return rec / div, total / div, kl_weight * km
