"""GPU parity of the CUDA RNN-VAE path (through the C-ABI) against the CPU oracle and the reference-generated goldens.

Stated tolerances (fp32 path evaluated with 3-pass bf16-split tensor-core products, fp32 accumulation):
  outputs (pred, future, z, mu, logvar, latent vectors)   <= 1e-4  max-norm relative   (north_star: 1e-4)
  loss terms                                                <= 2e-5  relative (KL: 1e-4, it is a difference of O(1) terms)
  gradients                                                 <= 1e-4  max-norm relative per tensor
The reference's own fp32-vs-fp64 noise floor is ~6e-8 on losses and ~6e-7 on gradients."""
import os

import numpy as np
import pytest
import torch

from oracle import vame_oracle as vo

pytestmark = pytest.mark.gpu

OUT_TOL, LOSS_TOL, KL_TOL, GRAD_TOL = 1e-4, 2e-5, 1e-4, 1e-4


def rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def make(T, Z, F, fut, S, H, seed=19):
    from vame_b200.engine import Engine
    torch.manual_seed(seed)
    port = vo.RefPort(2 * T, Z, F, bool(fut), S, hidden=H)
    eng = Engine(F, T, Z, H, H, H, bool(fut), S, False, device="cuda")
    eng.load_state_dict(port.state_dict())
    return port, eng


def check_step(port, eng, x, xf, eps, hp, golden=None):
    fut = port.future
    terms, grads, aux = vo.train_step(port, x, xf, eps, hp)
    out = eng.forward(x.cuda(), eps.cuda(), save=True)
    for k in ("pred", "future", "z", "mu", "logvar"):
        if k in out:
            assert rel(out[k], aux[k]) <= OUT_TOL, (k, rel(out[k], aux[k]))
            if golden is not None:
                assert rel(out[k], golden[k]) <= OUT_TOL, ("golden", k)
    cfg = eng.loss_cfg(kmeans_loss=hp["kmeans_loss"], kmeans_lambda=hp["kmeans_lambda"], bsize=hp["bsize"], beta=hp["beta"],
                       kl_weight=hp["kl_weight"])
    ls = eng.loss(cfg, xf.cuda() if fut else None, want_grads=True).cpu().tolist()
    ref = [terms["rec"], terms.get("fut", 0.0), terms["kl"], terms["kmeans"], terms["total"]]
    for i, (name, tol) in enumerate((("rec", LOSS_TOL), ("fut", LOSS_TOL), ("kl", KL_TOL), ("kmeans", LOSS_TOL), ("total", LOSS_TOL))):
        assert abs(ls[i] - ref[i]) <= tol * max(abs(ref[i]), 1e-3), (name, ls[i], ref[i])
    eng.backward(cfg)
    torch.cuda.synchronize()
    gv = eng.views(eng.grad)
    for k in eng.names:
        assert rel(gv[k], grads[k]) <= GRAD_TOL, (k, rel(gv[k], grads[k]))
    return terms, grads, aux


@pytest.mark.parametrize("name", ["tiny_fut", "small_nofut", "odd_fut", "c2_h256"])
def test_train_step_matches_reference_goldens(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "step_%s.npz" % name))
    B, T, F, Z, H, fut, S = (int(v) for v in g["cfg"])
    port, eng = make(T, Z, F, fut, S, H)
    x, xf, eps = (torch.from_numpy(g[k]) for k in ("x", "fut", "eps"))
    hp = dict(beta=1.0, kl_weight=float(g["hp_kl_weight"]), kmeans_loss=Z, kmeans_lambda=0.1, bsize=B)
    check_step(port, eng, x, xf, eps, hp, golden=g)
    # reference-generated gradients (not only the oracle's)
    gv = eng.views(eng.grad)
    for k in eng.names:
        if "grad/" + k in g:
            assert rel(gv[k], g["grad/" + k]) <= GRAD_TOL, k
        else:
            got = gv[k].reshape(-1)[::97][:512].cpu()
            assert float((got - torch.from_numpy(g["gsample/" + k])).abs().max()) <= GRAD_TOL * float(g["gsum/" + k][2]), k
    # eval mode: z = mu
    oe = eng.forward(x.cuda(), None, save=False)
    assert rel(oe["pred"], g["pred_eval"]) <= OUT_TOL
    assert torch.equal(oe["z"], oe["mu"])


@pytest.mark.parametrize("cfg", [
    dict(B=256, T=30, F=24, Z=30, S=0, fut=False),     # BASELINE configs[1]
    dict(B=256, T=30, F=24, Z=30, S=15, fut=True),
    dict(B=512, T=60, F=60, Z=50, S=30, fut=True),     # BASELINE configs[2]
    dict(B=200, T=30, F=12, Z=30, S=15, fut=True),     # ragged batch (not a multiple of the 128-row tile)
])
def test_train_step_full_size_vs_oracle(cfg):
    port, eng = make(cfg["T"], cfg["Z"], cfg["F"], cfg["fut"], cfg["S"], 256)
    x, xf, eps = vo.synthetic_batch(cfg["B"], cfg["T"], cfg["F"], max(cfg["S"], 1), cfg["Z"])
    hp = dict(beta=1.0, kl_weight=1.0, kmeans_loss=cfg["Z"], kmeans_lambda=0.1, bsize=cfg["B"])
    check_step(port, eng, x, xf[:, :cfg["S"]] if cfg["fut"] else xf, eps, hp)


def test_option_branches_vs_reference_golden(golden_dir):
    """Branches the default configuration never takes (VERDICT r1 "missing" 6): Lambda softplus (rnn_model.py:59-61,67), MSE
    reduction 'mean' for both reconstruction terms (rnn_vae.py:35-43), kmeans_loss < zdims incl. its gradient (rnn_vae.py:48),
    hidden_size_layer_1 != hidden_size_rec != hidden_size_pred - against the step the reference's own objects computed and
    against the oracle."""
    from vame_b200.engine import Engine
    g = np.load(os.path.join(golden_dir, "step_opts.npz"))
    B, T, F, Z, S, H1, HR, HP, K = (int(v) for v in g["cfg"])
    torch.manual_seed(19)
    port = vo.RefPort(2 * T, Z, F, True, S, hidden=H1, softplus=True, hidden_rec=HR, hidden_pred=HP)
    eng = Engine(F, T, Z, H1, HR, HP, True, S, True, device="cuda")
    eng.load_state_dict(port.state_dict())
    x, xf, eps = (torch.from_numpy(g[k]) for k in ("x", "fut", "eps"))
    klw, lam = float(g["hp_kl_weight"]), float(g["hp_lambda"])
    hp = dict(beta=1.0, kl_weight=klw, kmeans_loss=K, kmeans_lambda=lam, bsize=B, mse_red="mean", mse_pred="mean")
    terms, grads, aux = vo.train_step(port, x, xf, eps, hp)
    out = eng.forward(x.cuda(), eps.cuda(), save=True)
    for k in ("pred", "future", "z", "mu", "logvar"):
        assert rel(out[k], g[k]) <= OUT_TOL and rel(out[k], aux[k]) <= OUT_TOL, k
    assert float(out["logvar"].min()) >= 0.0
    cfg = eng.loss_cfg("mean", "mean", K, lam, B, 1.0, klw)
    ls = eng.loss(cfg, xf.cuda(), want_grads=True).cpu().tolist()
    for i, (name, tol) in enumerate((("rec", LOSS_TOL), ("fut", LOSS_TOL), ("kl", KL_TOL), ("kmeans", LOSS_TOL), ("total", LOSS_TOL))):
        ref = float(g["loss_" + name])
        assert abs(ls[i] - ref) <= tol * max(abs(ref), 1e-3), (name, ls[i], ref)
    eng.backward(cfg)
    gv = eng.views(eng.grad)
    for k in eng.names:
        assert rel(gv[k], g["grad/" + k]) <= GRAD_TOL and rel(gv[k], grads[k]) <= GRAD_TOL, (k, rel(gv[k], g["grad/" + k]))
    oe = eng.forward(x.cuda(), None, save=False)
    assert rel(oe["pred"], g["pred_eval"]) <= OUT_TOL


def test_amsgrad_three_steps_vs_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "step_tiny_fut.npz"))
    B, T, F, Z, H, fut, S = (int(v) for v in g["cfg"])
    port, eng = make(T, Z, F, fut, S, H)
    x, xf, eps = (torch.from_numpy(g[k]).cuda() for k in ("x", "fut", "eps"))
    cfg = eng.loss_cfg(kmeans_loss=Z, kmeans_lambda=0.1, bsize=B, beta=1.0, kl_weight=float(g["hp_kl_weight"]))
    for _ in range(3):
        eng.forward(x, eps, save=True, want=())
        eng.loss(cfg, xf, want_grads=True)
        eng.backward(cfg)
        eng.adam_step(lr=5e-4)
    v = eng.views()
    # Adam's first steps move each weight by ~lr*sign(g); elements whose gradient sits at the fp32 noise floor may land
    # elsewhere, hence a (tiny) outlier budget bounded by 3*lr (same criterion as tests/test_oracle_pinned.py)
    for k in eng.names:
        err = (v[k].cpu() - torch.from_numpy(g["w3/" + k])).abs()
        assert float((err > 5e-6).float().mean()) <= 2e-3 and float(err.max()) <= 3 * 5e-4 + 1e-6, k
    assert int(eng.opt_state["step"].item()) == 3


def test_submodule_forwards(golden_dir):
    g = np.load(os.path.join(golden_dir, "step_odd_fut.npz"))
    B, T, F, Z, H, fut, S = (int(v) for v in g["cfg"])
    port, eng = make(T, Z, F, fut, S, H)
    x, eps = torch.from_numpy(g["x"]), torch.from_numpy(g["eps"])
    with torch.no_grad():
        hid = port.encode(x)
        zr, mur, lvr = port.lmbda(hid, eps)
        pr = port.decode(zr, "decoder")
        pf = port.decode(zr, "decoder_future")
    assert rel(eng.encoder_forward(x.cuda()), hid) <= OUT_TOL
    z, mu, lv = eng.lambda_forward(hid.cuda(), eps.cuda())
    assert max(rel(z, zr), rel(mu, mur), rel(lv, lvr)) <= OUT_TOL
    assert rel(eng.decoder_forward(zr.cuda(), 0), pr) <= OUT_TOL
    assert rel(eng.decoder_forward(zr.cuda(), 1), pf) <= OUT_TOL


def test_cluster_loss_kernel():
    from vame_b200.engine import Engine
    from oracle import gru_numpy as gnp
    eng = Engine(12, 30, 30, device="cuda")
    rng = np.random.default_rng(0)
    for B, Z, k in ((16, 10, 10), (256, 30, 30), (8, 30, 30), (40, 13, 5), (4096, 50, 50), (300, 64, 64)):
        Lm = rng.standard_normal((B, Z)).astype(np.float32)
        loss, dl = eng.cluster_loss(torch.from_numpy(Lm).cuda(), k, 0.1, B, want_grad=True)
        ref, dref = gnp.cluster_loss(Lm.astype(np.float64), k, 0.1, B, return_grad=True)
        assert abs(loss.item() - ref) <= 1e-6 * abs(ref), (B, Z, k)
        assert rel(dl, dref) <= 1e-5, (B, Z, k)          # also for k < min(B, Z): only the leading singular values carry gradient


def test_embed_matches_reference_goldens(golden_dir):
    g = np.load(os.path.join(golden_dir, "embed_synth.npz"))
    T, F, Z, H = (int(v) for v in g["cfg"])
    port, eng = make(T, Z, F, True, 15, H)
    series = torch.from_numpy(np.ascontiguousarray(g["series"].T)).float().cuda()
    lat = eng.embed(series, chunk=256)
    assert tuple(lat.shape) == g["latent"].shape                       # (N - T, Z)
    assert rel(lat, g["latent"]) <= OUT_TOL
    assert rel(eng.embed(series, chunk=128), lat) <= 1e-6              # chunking does not change results
    # sub-ranges (what a data-parallel rank computes)
    part = eng.embed(series, first_window=100, n_windows=77, chunk=128)
    assert torch.equal(part, eng.embed(series, chunk=128)[100:177])
    assert eng.embed(series, first_window=0, n_windows=0).shape[0] == 0


def test_embed_video1(golden_dir):
    """BASELINE configs[0] data path: examples/video-1.csv -> reference csv_to_numpy/create_trainset -> latent vectors.
    Reference vectors (literal batch-1 loop) are committed for every 40th of the first 5970 windows; the remaining
    windows are checked against the oracle port."""
    g = np.load(os.path.join(golden_dir, "video1.npz"))
    T, F, Z, H = (int(v) for v in g["cfg"])
    port, eng = make(T, Z, F, True, 15, H)
    clean = g["clean"]                                                   # (F, N) float32
    series = torch.from_numpy(np.ascontiguousarray(clean.T)).cuda()
    lat = eng.embed(series, chunk=8192)
    assert lat.shape[0] == clean.shape[1] - T
    n = int(g["n_ref_windows"])
    assert rel(lat[:n][::40], g["latent_first"]) <= OUT_TOL
    ref = vo.embed_batched(port, clean.astype(np.float64), T)
    assert rel(lat, ref) <= OUT_TOL
    from vame_b200 import _lib
    assert _lib.lib().vame_get_option(b"rows_timeouts") == 0, "a bounded wait inside gru_rows_fwd_kernel gave up"


def test_module_surface_and_autograd(golden_dir):
    """nn.Module mirror: state_dict round trip, eval forward, and loss.backward() through the whole-model autograd node
    with the reference-style loss composition (rnn_vae.py:124-129)."""
    from vame_b200.rnn_model import RNN_VAE
    from vame_b200 import rnn_vae as rv
    g = np.load(os.path.join(golden_dir, "step_small_nofut.npz"))
    B, T, F, Z, H, fut, S = (int(v) for v in g["cfg"])
    torch.manual_seed(19)
    model = RNN_VAE(2 * T, Z, F, bool(fut), S, H, H, H, H, 0, 0, 0, False).cuda()
    torch.manual_seed(19)
    port = vo.RefPort(2 * T, Z, F, bool(fut), S, hidden=H)
    sd = model.state_dict()
    for k, v in port.state_dict().items():
        assert torch.equal(sd[k].cpu(), v), k
    x, eps = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["eps"]).cuda()
    model.eval()
    with torch.no_grad():
        pred, z, mu, lv = model(x)
    assert rel(pred, g["pred_eval"]) <= OUT_TOL
    # training-mode autograd with injected eps
    model.train()
    orig = torch.randn
    torch.randn = lambda *a, **k: eps if tuple(a[:2]) == (B, Z) else orig(*a, **k)
    try:
        pred, z, mu, lv = model(x)
    finally:
        torch.randn = orig
    klw = float(g["hp_kl_weight"])
    loss = rv.reconstruction_loss(x, pred, "sum") + klw * rv.kullback_leibler_loss(mu, lv) + klw * rv.cluster_loss(z.T, Z, 0.1, B)
    assert abs(loss.item() - float(g["loss_total"])) <= LOSS_TOL * abs(float(g["loss_total"]))
    loss.backward()
    for k, p in model.named_parameters():
        assert rel(p.grad, g["grad/" + k]) <= GRAD_TOL, k
    # a torch optimizer stepping the (view) parameters must be picked up by the next forward: optimizer.step() bumps the
    # PARAMETERS' version counters, not the flat buffer's - compare eval outputs with the oracle after the
    # same update, before and after
    model.eval()
    with torch.no_grad():
        before = model(x)[0].clone()
    opt = torch.optim.Adam(model.parameters(), lr=5e-3, amsgrad=True)
    opt.step()
    with torch.no_grad():
        after = model(x)[0].clone()
    port.load_state_dict({k: v.cpu() for k, v in model.state_dict().items()})       # the updated weights
    with torch.no_grad():
        ref_after = port.forward(x.cpu(), None)[0]
    assert rel(after, ref_after) <= OUT_TOL, "forward after optimizer.step() used stale packed weights"
    assert rel(before, ref_after) > 1e-3                                            # the update was not a no-op
    # ... and so must a no-grad in-place edit of a parameter
    with torch.no_grad():
        model.decoder.hidden_to_output.bias.add_(0.25)
        shifted = model(x)[0]
    assert rel(shifted, ref_after + 0.25) <= OUT_TOL
    # two training forwards of the same batch size before one backward: the first node's activations are gone -> loud error
    model.train()
    p_a = model(x)[0]
    p_b = model(x)[0]
    p_b.sum().backward()
    with pytest.raises(RuntimeError, match="overwritten"):
        p_a.sum().backward()
    # an evaluation forward in between does not disturb the saved activations
    model.zero_grad()
    p_c = model(x)[0]
    model.eval()
    with torch.no_grad():
        model(x)
    model.train()
    p_c.sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())


def test_noisy_input_clean_target(golden_dir):
    """cfg['noise']: the model sees x + noise, the reconstruction loss compares against the clean x (rnn_vae.py:116-124)."""
    g = np.load(os.path.join(golden_dir, "step_small_nofut.npz"))
    B, T, F, Z, H, fut, S = (int(v) for v in g["cfg"])
    port, eng = make(T, Z, F, fut, S, H)
    x, eps = torch.from_numpy(g["x"]), torch.from_numpy(g["eps"])
    torch.manual_seed(3)
    xn = x + 0.3 * torch.randn_like(x)
    port.zero_grad()
    pred, z, mu, lv = port.forward(xn, eps)
    ref = vo.reconstruction_loss(x, pred, "sum") + 0.7 * vo.kullback_leibler_loss(mu, lv) + 0.7 * vo.cluster_loss(z.T, Z, 0.1, B)
    ref.backward()
    eng.forward(xn.cuda(), eps.cuda(), save=True, want=())
    cfg = eng.loss_cfg(kmeans_loss=Z, kmeans_lambda=0.1, bsize=B, beta=1.0, kl_weight=0.7)
    ls = eng.loss(cfg, None, want_grads=True, target=x.cuda()).cpu().tolist()
    assert abs(ls[4] - ref.item()) <= LOSS_TOL * abs(ref.item())
    eng.backward(cfg)
    gv = eng.views(eng.grad)
    for k, p in port.named_parameters():
        assert rel(gv[k], p.grad) <= GRAD_TOL, k


def _sampler_project(tmp_path, F=6, N=3000, seed=0):
    rng = np.random.default_rng(seed)
    series = np.cumsum(rng.standard_normal((F, N)), axis=1)
    d = str(tmp_path) + os.sep
    np.save(d + "train_seq.npy", series)
    np.save(d + "test_seq.npy", series[:, :600])
    return d, series


def test_sample_windows_kernel_bit_exact_and_seeded(tmp_path):
    """csrc/sampler.cu vs the reference dataset arithmetic (vame/model/dataloader.py:45-56 + rnn_vae.py:107-112): for the starts
    the kernel reports, data / future are BIT-IDENTICAL to ((X[:, s:s+2T] - mean) / std) in float64 -> float32; starts are
    uniform in [0, N - 2T); the stream is reproducible from (seed, draw counter) and advances on the device."""
    from vame_b200.dataloader import DeviceWindowSampler
    T, S, Z, B = 10, 4, 5, 512
    d, series = _sampler_project(tmp_path)
    sm = DeviceWindowSampler.from_files(d, "train_seq.npy", True, 2 * T, B, seed=7)
    F, N = series.shape
    mean, std = series.mean(), series.std()
    x = torch.empty(B, T, F, device="cuda")
    fut = torch.empty(B, S, F, device="cuda")
    eps = torch.empty(B, Z, device="cuda")
    starts = torch.empty(B, dtype=torch.int64, device="cuda")
    sm.fill(x, fut, eps, starts_out=starts)
    st = starts.cpu().numpy()
    assert st.min() >= 0 and st.max() < N - 2 * T and len(np.unique(st)) > B // 2
    ref = np.stack([((series[:, s:s + 2 * T] - mean) / std).T.astype(np.float32) for s in st])          # (B, 2T, F)
    assert np.array_equal(x.cpu().numpy(), ref[:, :T]) and np.array_equal(fut.cpu().numpy(), ref[:, T:T + S])
    assert int(sm.counter.item()) == 1
    # explicit starts (edge windows) reproduce the same arithmetic; a second draw differs; the same seed repeats the stream
    edge = torch.tensor([0, N - 2 * T - 1] * (B // 2), dtype=torch.int64, device="cuda")
    x2 = torch.empty_like(x)
    sm.fill(x2, starts=edge)
    assert np.array_equal(x2[1].cpu().numpy(), ((series[:, N - 2 * T - 1:N - 1] - mean) / std).T.astype(np.float32)[:T])
    sm2 = DeviceWindowSampler(sm.dataset, B, seed=7)
    xa, ea = torch.empty_like(x), torch.empty_like(eps)
    sm2.fill(xa, None, ea)
    assert torch.equal(xa, x) and torch.equal(ea, eps)
    sm2.fill(xa, None, ea)
    assert not torch.equal(xa, x) and not torch.equal(ea, eps)
    # distribution: starts ~ U[0, N - 2T), eps ~ N(0, 1)
    big = torch.empty(8192, Z, device="cuda")
    xs = torch.empty(8192, T, F, device="cuda")
    so = torch.empty(8192, dtype=torch.int64, device="cuda")
    sm.fill(xs, None, big, starts_out=so)
    u = so.double().cpu().numpy() / (N - 2 * T)
    assert abs(u.mean() - 0.5) < 0.02 and abs(u.var() - 1 / 12) < 0.01
    e = big.double().cpu().numpy().ravel()
    assert abs(e.mean()) < 0.02 and abs(e.std() - 1) < 0.02 and abs((e ** 4).mean() - 3) < 0.15


def test_device_window_sampler_and_train_epoch(tmp_path):
    """the dataset / loader pair that install() binds into vame.model.rnn_vae feeds the drop-in train()/test();
    statistics files and batch shapes follow the reference dataset; with the sampler captured in the step graph an epoch does no
    host work per batch and the loss goes down over a few epochs."""
    from vame_b200 import _lib
    from vame_b200.dataloader import SEQUENCE_DATASET, Data, DeviceWindowSampler
    from vame_b200.rnn_model import RNN_VAE
    from vame_b200 import rnn_vae as rv
    T, F, Z, H, S, B = 10, 6, 5, 32, 4, 64
    d, series = _sampler_project(tmp_path, F=F)
    trainset = SEQUENCE_DATASET(d, data="train_seq.npy", train=True, temporal_window=2 * T)          # rnn_vae.py:326-330
    testset = SEQUENCE_DATASET(d, data="test_seq.npy", train=False, temporal_window=2 * T)
    tr = Data.DataLoader(trainset, batch_size=B, shuffle=True, drop_last=True)
    te = Data.DataLoader(testset, batch_size=B // 4, shuffle=True, drop_last=True)
    assert isinstance(tr, DeviceWindowSampler) and isinstance(te, DeviceWindowSampler)
    assert abs(float(np.load(d + "seq_mean.npy")) - series.mean()) < 1e-12 and len(tr) == 3000 // B and len(trainset) == 3000
    item = trainset[0]
    assert tuple(item.shape) == (F, 2 * T) and item.dtype == torch.float64
    batch = next(iter(tr))
    assert tuple(batch.shape) == (B, F, 2 * T) and batch.is_cuda
    torch.manual_seed(19)
    model = RNN_VAE(2 * T, Z, F, True, S, H, H, H, H, 0, 0, 0, False).cuda()
    opt = torch.optim.Adam(model.parameters(), lr=5e-3, amsgrad=True)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=100, gamma=1)
    first = last = None
    lib = _lib.lib()
    for epoch in range(1, 5):
        n0 = lib.vame_launch_count()
        r = rv.train(tr, epoch, model, opt, "linear", 1, 0, 4, 2 * T, True, S, sched, "sum", "sum", Z, 0.1, B, epoch == 4)
        if epoch == 3:            # steady state (graph captured in epoch 1): replays only, the host launches nothing per batch
            assert lib.vame_launch_count() == n0
        first = r[4] if first is None else first
        last = r[4]
        assert all(np.isfinite(v) for v in r)
    assert last < 0.8 * first, (first, last)
    mse, loss, km = rv.test(te, 4, model, opt, 1, 1.0, 2 * T, "sum", Z, 0.1, True, B // 4)
    assert np.isfinite(mse) and np.isfinite(loss)
