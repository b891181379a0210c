"""GPU parity of the drop-in train()/test() epoch functions with the reference's own train()/test()
(fixture produced by oracle/gen_golden.py: 5 fixed batches, future decoder, StepLR).  Lambda's eps is drawn from the
RNG in both implementations, so the comparison injects the reference's CPU draws."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_train_and_test_epoch_match_reference(golden_dir):
    from vame_b200.rnn_model import RNN_VAE
    from vame_b200 import rnn_vae as rv
    g = np.load(os.path.join(golden_dir, "train_fn.npz"))
    B, T, F, Z, H, fut, S = (int(v) for v in g["cfg"])
    torch.manual_seed(19)
    model = RNN_VAE(2 * T, Z, F, True, S, H, H, H, H, 0, 0, 0, False).cuda()
    batches = [torch.from_numpy(b) for b in g["batches"]]
    opt = torch.optim.Adam(model.parameters(), lr=5e-4, amsgrad=True)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=100, gamma=1)
    # the reference drew eps with torch.manual_seed(123) then randn_like on CPU, one (B, Z) draw per batch
    torch.manual_seed(123)
    draws = [torch.randn(B, Z) for _ in batches]
    it = iter(draws)
    orig = torch.randn
    torch.randn = lambda *a, **k: next(it).to(k.get("device", "cpu")) if tuple(a[:2]) == (B, Z) else orig(*a, **k)
    try:
        ret = rv.train(batches, 3, model, opt, "linear", 1, 0, 4, 2 * T, True, S, sched, "sum", "sum", Z, 0.1, B, False)
    finally:
        torch.randn = orig
    ref = g["train_ret"]
    assert ret[0] == ref[0]
    for a, b in zip(ret[1:], ref[1:]):
        assert abs(a - b) <= 2e-4 * max(abs(b), 1e-3), (ret, ref)
    sd = model.state_dict()
    for k in sd:
        err = (sd[k].cpu() - torch.from_numpy(g["w_after/" + k])).abs()
        assert float((err > 1e-5).float().mean()) <= 1e-3 and float(err.max()) <= 5 * 5e-4 + 1e-6, k
    ret_t = rv.test(batches, 3, model, opt, 1, ret[0], 2 * T, "sum", Z, 0.1, True, B)
    for a, b in zip(ret_t, g["test_ret"]):
        assert abs(float(a) - b) <= 5e-3 * max(abs(b), 1e-3), (ret_t, g["test_ret"])


def test_trainstep_staged_host_batches_match_device_batches():
    """TrainStep.load() with pinned HOST tensors (copy stream + staging sets, uploads issued one step ahead) must give the same
    losses and weights as device-resident batches."""
    from oracle import vame_oracle as vo
    from vame_b200.engine import Engine, TrainStep
    T, Z, F, H, B = 12, 8, 10, 64, 48
    torch.manual_seed(19)
    port = vo.RefPort(2 * T, Z, F, False, 0, hidden=H)
    batches = [vo.synthetic_batch(B, T, F, 1, Z, seed=30 + i) for i in range(4)]
    results = []
    for staged in (False, True):
        eng = Engine(F, T, Z, H, H, H, False, 0, False, device="cuda")
        eng.load_state_dict(port.state_dict())
        cfg = eng.loss_cfg(kmeans_loss=Z, kmeans_lambda=0.1, bsize=B, beta=1.0, kl_weight=1.0)
        eng.set_hyper(lr=5e-4, kl_weight=1.0, beta=1.0, kmeans_lambda=0.1)
        ts = TrainStep(eng, B, cfg, world=1)
        ts.capture()
        losses = []
        if staged:
            host = [(x.pin_memory(), eps.pin_memory()) for x, _, eps in batches]
            ts.load(host[0][0], None, host[0][1])
            for i in range(len(host)):
                out = ts.run()
                if i + 1 < len(host):
                    ts.load(host[i + 1][0], None, host[i + 1][1])      # uploaded while step i computes
                losses.append(out.clone())
        else:
            for x, _, eps in batches:
                ts.load(x.cuda(), None, eps.cuda())
                losses.append(ts.run().clone())
        torch.cuda.synchronize()
        results.append((torch.stack(losses).cpu(), eng.flat.detach().cpu().clone()))
    # (split-K accumulation uses red.add, so two runs agree to rounding, not bit for bit)
    assert torch.allclose(results[0][0], results[1][0], rtol=1e-5, atol=1e-5)
    # AMSGrad moves a weight by ~lr whatever the size of its gradient, so elements with near-zero gradients amplify rounding
    # differences: same criterion as the reference comparison above (almost all weights equal, none off by more than 4 steps of lr)
    err = (results[0][1] - results[1][1]).abs()
    assert float((err > 1e-5).float().mean()) <= 1e-3 and float(err.max()) <= 4 * 5e-4 + 1e-6


def test_trainstep_product_path_vs_oracle_and_repack_for_other_batch_sizes():
    """The train loop's own path (TrainStep: k-means prior armed in the forward pass, one graph, partial weight re-pack) against
    the oracle for two optimizer steps at H = 256, three 128-row tiles (ragged 300-window batch -> two waves of resident-weight
    clusters); then a forward pass at a batch size that runs on the slice kernels (B = 600), whose weight formats the train
    loop skipped - the engine must re-pack them from the updated parameters."""
    from oracle import vame_oracle as vo
    from vame_b200.engine import Engine, TrainStep
    T, Z, F, H, B = 5, 30, 24, 256, 300
    torch.manual_seed(19)
    port = vo.RefPort(2 * T, Z, F, False, 0, hidden=H)
    eng = Engine(F, T, Z, H, H, H, False, 0, False, device="cuda")
    eng.load_state_dict(port.state_dict())
    opt = vo.make_optimizer(port)
    hp = dict(beta=1.0, kl_weight=0.7, kmeans_loss=Z, kmeans_lambda=0.1, bsize=B)
    cfg = eng.loss_cfg(kmeans_loss=Z, kmeans_lambda=0.1, bsize=B, beta=1.0, kl_weight=0.7)
    eng.set_hyper(lr=5e-4, kl_weight=0.7, beta=1.0, kmeans_lambda=0.1)
    ts = TrainStep(eng, B, cfg, world=1)
    assert ts.capture()
    for i in range(2):
        x, xf, eps = vo.synthetic_batch(B, T, F, 1, Z, seed=40 + i)
        terms, grads, _ = vo.train_step(port, x, xf, eps, hp, optimizer=opt)
        ts.load(x.cuda(), None, eps.cuda())
        ls = ts.run().cpu().tolist()
        tol = 5e-5 if i == 0 else 3e-4        # (step 2 starts from weights that differ by AMSGrad's rounding sensitivity)
        assert abs(ls[4] - terms["total"]) <= tol * abs(terms["total"]), (i, ls, terms)
        gv = eng.views(eng.grad)
        if i == 0:                    # (after the first update the two weight sets differ by AMSGrad's rounding sensitivity)
            for k in eng.names:
                e = float((gv[k].double().cpu() - grads[k].double()).abs().max() / grads[k].double().abs().max().clamp_min(1e-30))
                assert e <= 1e-4, (k, e)
    # a batch size served by the slice kernels: needs the W_hh formats the train loop did not refresh
    xb, _, _ = vo.synthetic_batch(600, T, F, 1, Z, seed=77)
    out = eng.forward(xb.cuda(), None, save=False, want=("mu",))
    twin = vo.RefPort(2 * T, Z, F, False, 0, hidden=H)
    twin.load_state_dict({k: v.cpu() for k, v in eng.state_dict().items()})
    with torch.no_grad():
        mu_ref = twin.lmbda(twin.encode(xb), None)[1]
    e = float((out["mu"].double().cpu() - mu_ref.double()).abs().max() / mu_ref.double().abs().max())
    assert e <= 1e-4, e


def test_load_model_and_embedd_latent_vectors_trained_checkpoint(tmp_path, golden_dir):
    """The two inference drop-ins on a real project layout (vame/analysis/pose_segmentation.py:27-64, 67-101): a `.pkl` written
    in the reference's format (torch.save of the state_dict, rnn_vae.py:367) holding the weights the REFERENCE's own train()
    produced in 10 epochs on video-1 (oracle/gen_golden.py:gen_video1_trained; split-precision error grows with |W_hh|, which
    initial weights do not exercise), `<file>-PE-seq-clean.npy` as create_trainset leaves it; the latent vectors must match the
    reference's literal batch-1 loop (every 40th of the first 5970 windows) within 1e-4 and the oracle on all 29 967."""
    from collections import OrderedDict
    from oracle import vame_oracle as vo
    from vame_b200 import pose_segmentation as ps
    g = np.load(os.path.join(golden_dir, "video1_trained.npz"))
    clean = np.load(os.path.join(golden_dir, "video1.npz"))["clean"].astype(np.float64)      # (F, N)
    T, F, Z, H = (int(v) for v in g["cfg"])
    proj = tmp_path / "proj"
    (proj / "model" / "best_model").mkdir(parents=True)
    (proj / "data" / "video-1").mkdir(parents=True)
    np.save(proj / "data" / "video-1" / "video-1-PE-seq-clean.npy", clean)
    sd = OrderedDict((k[2:], torch.from_numpy(g[k])) for k in g.files if k.startswith("w/"))
    torch.save(sd, proj / "model" / "best_model" / "VAME_proj.pkl")
    cfg = dict(project_path=str(proj), Project="proj", time_window=T, num_features=F, zdims=Z, prediction_decoder=0,
               prediction_steps=15, hidden_size_layer_1=H, hidden_size_layer_2=H, hidden_size_rec=H, hidden_size_pred=H,
               dropout_encoder=0, dropout_rec=0, dropout_pred=0, softplus=False)
    model = ps.load_model(cfg, "VAME", True)
    assert not model.training and list(model.state_dict().keys()) == list(sd.keys())
    whh = max(float(v.abs().max()) for k, v in sd.items() if "weight_hh" in k)
    assert whh > 1.5 / np.sqrt(H), "the fixture is supposed to hold TRAINED recurrent weights"
    lat = ps.embedd_latent_vectors(cfg, ["video-1"], model, True)
    assert len(lat) == 1 and lat[0].shape == (clean.shape[1] - T, Z) and lat[0].dtype == np.float32
    n = int(g["n_ref_windows"])
    ref = g["latent_first"]
    err = np.abs(lat[0][:n][::40] - ref).max() / np.abs(ref).max()
    assert err <= 1e-4, err
    port = vo.RefPort(2 * T, Z, F, False, 0, hidden=H).load_state_dict(sd)
    full = vo.embed_batched(port, clean, T)
    err_full = np.abs(lat[0] - full).max() / np.abs(full).max()
    assert err_full <= 1e-4, err_full
