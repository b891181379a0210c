"""Pin the oracle (oracle/vame_oracle.py torch port and oracle/gru_numpy.py numpy restatement)
against fixtures produced by the UNMODIFIED reference (oracle/gen_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import gru_numpy as gnp
from oracle import vame_oracle as vo
from oracle.ref_shim import reference_available

STEP_CASES = ["tiny_fut", "small_nofut", "odd_fut", "c2_h256"]


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _port_for(g, seed=19):
    B, T, F, Z, H, fut, S = (int(v) for v in g["cfg"])
    torch.manual_seed(seed)
    return vo.RefPort(2 * T, Z, F, bool(fut), S, hidden=H), (B, T, F, Z, H, bool(fut), S)


def _hp(g, Z, B, fut, S):
    return dict(beta=1.0, kl_weight=float(g["hp_kl_weight"]), kmeans_loss=Z, kmeans_lambda=0.1, bsize=B,
                future=fut, steps_future=S)


@pytest.mark.parametrize("name", STEP_CASES)
def test_port_matches_reference_step(golden_dir, name):
    g = _load(golden_dir, "step_%s.npz" % name)
    port, (B, T, F, Z, H, fut, S) = _port_for(g)
    if "w/encoder.encoder_rnn.weight_hh_l0" in g:            # same seed -> same init as the reference
        for k, p in port.named_parameters():
            np.testing.assert_array_equal(p.detach().numpy(), g["w/" + k])
    x, xf, eps = (torch.from_numpy(g[k]) for k in ("x", "fut", "eps"))
    terms, grads, aux = vo.train_step(port, x, xf, eps, _hp(g, Z, B, fut, S))
    for k in ("rec", "kl", "kmeans", "total") + (("fut",) if fut else ()):
        assert abs(terms[k] - float(g["loss_" + k])) <= 2e-6 * max(1.0, abs(float(g["loss_" + k]))), k
    for k in ("pred", "z", "mu", "logvar"):
        np.testing.assert_allclose(aux[k].numpy(), g[k], rtol=1e-5, atol=1e-6)
    for k, gr in grads.items():
        gr = gr.numpy()
        if "grad/" + k in g:
            ref = g["grad/" + k]
            assert np.abs(gr - ref).max() <= 2e-5 * max(np.abs(ref).max(), 1e-6), k
        else:
            ref = g["gsample/" + k]
            got = gr.reshape(-1)[::97][:512]
            assert np.abs(got - ref).max() <= 2e-5 * max(float(g["gsum/" + k][2]), 1e-6), k


@pytest.mark.parametrize("name", ["tiny_fut", "small_nofut", "odd_fut"])
def test_numpy_restatement_matches_reference_step(golden_dir, name):
    """fp64 numpy forward + hand-written BPTT vs the reference's fp32 autograd."""
    g = _load(golden_dir, "step_%s.npz" % name)
    B, T, F, Z, H, fut, S = (int(v) for v in g["cfg"])
    w = {k[2:]: g[k].astype(np.float64) for k in g.files if k.startswith("w/")}
    hp = _hp(g, Z, B, bool(fut), S)
    losses, grads, aux = gnp.train_step(w, g["x"].astype(np.float64), g["fut"].astype(np.float64),
                                        g["eps"].astype(np.float64), hp)
    for k in ("rec", "kl", "kmeans", "total") + (("fut",) if fut else ()):
        assert abs(losses[k] - float(g["loss_" + k])) <= 3e-6 * max(1.0, abs(float(g["loss_" + k]))), k
    np.testing.assert_allclose(aux["pred"], g["pred"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(aux["mu"], g["mu"], rtol=2e-5, atol=2e-6)
    for k, gr in grads.items():
        ref = g["grad/" + k]
        assert gr.shape == ref.shape, k
        assert np.abs(gr - ref).max() <= 5e-5 * max(np.abs(ref).max(), 1e-6), (k, np.abs(gr - ref).max(), np.abs(ref).max())


@pytest.mark.parametrize("name", ["tiny_fut", "small_nofut", "odd_fut"])
def test_amsgrad_three_steps(golden_dir, name):
    g = _load(golden_dir, "step_%s.npz" % name)
    port, (B, T, F, Z, H, fut, S) = _port_for(g)
    x, xf, eps = (torch.from_numpy(g[k]) for k in ("x", "fut", "eps"))
    hp = _hp(g, Z, B, fut, S)
    # torch port + torch Adam
    opt = vo.make_optimizer(port)
    for _ in range(3):
        vo.train_step(port, x, xf, eps, hp, optimizer=opt)
    for k, p in port.named_parameters():
        np.testing.assert_allclose(p.detach().numpy(), g["w3/" + k], rtol=0, atol=2e-6)
    # numpy restatement of AMSGrad driven by numpy gradients (fp64)
    w = {k[2:]: g[k].astype(np.float64) for k in g.files if k.startswith("w/")}
    st = {k: [np.zeros_like(v), np.zeros_like(v), np.zeros_like(v)] for k, v in w.items()}
    for step in (1, 2, 3):
        _, grads, _ = gnp.train_step(w, g["x"].astype(np.float64), g["fut"].astype(np.float64),
                                     g["eps"].astype(np.float64), hp)
        for k in w:
            gnp.amsgrad_step(w[k], grads[k], st[k][0], st[k][1], st[k][2], step, 5e-4)
    # Adam's first steps move every weight by ~lr*sign(g): elements whose gradient is at the fp32 noise
    # floor can legitimately land elsewhere, so allow a 1e-4 fraction of outliers bounded by 3*lr.
    for k in w:
        err = np.abs(w[k] - g["w3/" + k])
        assert (err > 5e-6).mean() <= 1e-4 and err.max() <= 3 * 5e-4 + 1e-6, k


def test_cluster_loss_zxz_equals_bxb(golden_dir):
    rng = np.random.default_rng(0)
    for B, Z, k in ((16, 10, 10), (64, 30, 30), (8, 30, 30), (40, 12, 5)):
        L = rng.standard_normal((B, Z))
        a = gnp.cluster_loss(L, k, 0.1, B)
        b = gnp.cluster_loss_bxb_svd(L, k, 0.1, B)
        assert abs(a - b) < 1e-10 * max(1, abs(b))
        lt = torch.from_numpy(L).requires_grad_(True)
        c = vo.cluster_loss(lt.T, k, 0.1, B)
        assert abs(c.item() - b) < 1e-9
        if B >= Z and k == Z:
            c.backward()
            _, dl = gnp.cluster_loss(L, k, 0.1, B, return_grad=True)
            np.testing.assert_allclose(dl, lt.grad.numpy(), rtol=1e-7, atol=1e-10)


def test_embed_synth(golden_dir):
    g = _load(golden_dir, "embed_synth.npz")
    T, F, Z, H = (int(v) for v in g["cfg"])
    torch.manual_seed(19)
    port = vo.RefPort(2 * T, Z, F, True, 15, hidden=H)
    lat_loop = vo.embed_loop(port, g["series"], T, limit=40)
    np.testing.assert_allclose(lat_loop, g["latent"][:40], rtol=1e-5, atol=1e-6)
    lat_b = vo.embed_batched(port, g["series"], T)
    assert lat_b.shape == g["latent"].shape
    assert np.abs(lat_b - g["latent"]).max() <= 1e-5 * np.abs(g["latent"]).max()
    lat_np = gnp.embed_windows(port.numpy_weights(np.float64), g["series"], T)
    assert np.abs(lat_np - g["latent"]).max() <= 1e-5 * np.abs(g["latent"]).max()


def test_embed_video1_subsample(golden_dir):
    g = _load(golden_dir, "video1.npz")
    T, F, Z, H = (int(v) for v in g["cfg"])
    torch.manual_seed(19)
    port = vo.RefPort(2 * T, Z, F, True, 15, hidden=H)
    n = int(g["n_ref_windows"])
    lat = vo.embed_batched(port, g["clean"].astype(np.float64), T, limit=n)
    ref = g["latent_first"]
    assert np.abs(lat[::40] - ref).max() <= 1e-5 * np.abs(ref).max()


def test_port_option_branches_match_reference(golden_dir):
    """softplus, 'mean' MSE reductions, kmeans_loss < zdims and three different hidden sizes: the oracle port vs the step the
    reference's own objects computed (oracle/gen_golden.py:gen_step_opts)."""
    g = _load(golden_dir, "step_opts.npz")
    B, T, F, Z, S, H1, HR, HP, K = (int(v) for v in g["cfg"])
    torch.manual_seed(19)
    port = vo.RefPort(2 * T, Z, F, True, S, hidden=H1, softplus=True, hidden_rec=HR, hidden_pred=HP)
    for k, p in port.named_parameters():                         # same seed, same construction order -> same init
        np.testing.assert_array_equal(p.detach().numpy(), g["w/" + k])
    x, xf, eps = (torch.from_numpy(g[k]) for k in ("x", "fut", "eps"))
    hp = dict(beta=1.0, kl_weight=float(g["hp_kl_weight"]), kmeans_loss=K, kmeans_lambda=float(g["hp_lambda"]), bsize=B,
              mse_red="mean", mse_pred="mean")
    terms, grads, aux = vo.train_step(port, x, xf, eps, hp)
    for k in ("rec", "fut", "kl", "kmeans", "total"):
        assert abs(terms[k] - float(g["loss_" + k])) <= 2e-6 * max(1.0, abs(float(g["loss_" + k]))), k
    for k in ("pred", "future", "z", "mu", "logvar"):
        np.testing.assert_allclose(aux[k].numpy(), g[k], rtol=1e-5, atol=1e-6)
    assert float(aux["logvar"].min()) >= 0.0                     # softplus
    for k, gr in grads.items():
        ref = g["grad/" + k]
        assert np.abs(gr.numpy() - ref).max() <= 2e-5 * max(np.abs(ref).max(), 1e-6), k


@pytest.mark.reference
@pytest.mark.skipif(not reference_available(), reason="reference not mounted")
def test_port_vs_live_reference_forward():
    from oracle.ref_shim import reference_modules
    rm, rv, ps = reference_modules()
    torch.manual_seed(19)
    ref = rm.RNN_VAE(60, 30, 24, True, 15, 256, 256, 256, 256, 0, 0, 0, False)
    torch.manual_seed(19)
    port = vo.RefPort(60, 30, 24, True, 15, hidden=256)
    sd = ref.state_dict()
    for k, p in port.named_parameters():
        assert torch.equal(p.detach(), sd[k]), k
    assert list(sd.keys()) == [k for k, _ in port.named_parameters()]
    x, xf, eps = vo.synthetic_batch(16, 30, 24, 15, 30)
    ref.eval()
    with torch.no_grad():
        a = ref(x)
        b = port.forward(x, None)
    for u, v in zip(a, b):
        assert torch.allclose(u, v, rtol=1e-6, atol=1e-6)


# ---- k-means: numpy restatement vs the reference's own same/individual_parameterization outputs --------
def _kmeans_gold():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kmeans_blobs.npz"))


def test_kmeans_seeding_and_lloyd_match_sklearn():
    from oracle import kmeans_numpy as K
    g = _kmeans_gold()
    X, k = g["X"], int(g["k"][0])
    _, idx = K.kmeans_plusplus(X, k, np.random.RandomState(int(g["pp_seed"][0])))
    assert np.array_equal(idx, g["pp_idx"])
    lab, cen, inertia, n_iter = K.lloyd(X, g["pp_init"])
    assert np.array_equal(lab, g["single_labels"])
    assert np.abs(cen - g["single_centers"]).max() < 1e-5
    assert abs(inertia - float(g["single_inertia"][0])) / inertia < 1e-6
    assert n_iter == int(g["single_n_iter"][0])


def test_kmeans_full_fit_matches_reference_parameterizations():
    from oracle import kmeans_numpy as K
    g = _kmeans_gold()
    X, k, sp = g["X"], int(g["k"][0]), int(g["split"][0])
    for i, (a, b) in enumerate(((0, sp), (sp, X.shape[0]))):          # individual_parameterization: seed 42, n_init 3
        lab, cen, _, _ = K.kmeans(X[a:b], k, 42, 3)
        assert np.array_equal(lab, g["ind_labels%d" % i])
        assert np.abs(cen - g["ind_centers%d" % i]).max() < 1e-5
