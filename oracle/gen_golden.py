"""TEST INFRASTRUCTURE ONLY — generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the dev container (needs /root/reference):  ``python -m oracle.gen_golden``
The reference ships no golden vectors of its own, so these fixtures — outputs of the
reference's own code on seeded inputs — are what pins the oracle (and through it the CUDA
path).  Everything here goes through the reference's public classes / functions:
RNN_VAE, reconstruction_loss, future_reconstruction_loss, cluster_loss,
kullback_leibler_loss, train, embedd_latent_vectors, csv_to_numpy, create_trainset.
"""
import os
import sys
import tempfile

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_shim import reference_modules, load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: (B, T, F, Z, H, future, S)
    "tiny_fut": (6, 5, 4, 3, 32, True, 3),
    "small_nofut": (16, 8, 12, 10, 64, False, 0),
    "odd_fut": (13, 7, 10, 6, 32, True, 4),        # B not a multiple of anything
    "c2_h256": (32, 30, 24, 30, 256, True, 15),    # BASELINE shapes at reduced batch
}


def _inputs(B, T, F, S, Z, seed=19):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 2 * T, F, generator=g)
    eps = torch.randn(B, Z, generator=g)
    return x[:, :T].contiguous(), x[:, T:T + S].contiguous(), eps


def _ref_model(rm, T, Z, F, fut, S, H, seed=19):
    torch.manual_seed(seed)                                      # rnn_vae.py:292
    return rm.RNN_VAE(2 * T, Z, F, fut, S, H, H, H, H, 0, 0, 0, False)


def _with_eps(eps, fn):
    """Run fn() with torch.randn_like returning the injected eps (rnn_model.py:73)."""
    orig = torch.randn_like
    torch.randn_like = lambda t, *a, **k: eps.to(t.dtype)
    try:
        return fn()
    finally:
        torch.randn_like = orig


def gen_step_case(name, rm, rv):
    B, T, F, Z, H, fut, S = CASES[name]
    model = _ref_model(rm, T, Z, F, fut, S, H)
    x, xf, eps = _inputs(B, T, F, S, Z)
    hp = dict(beta=1.0, kl_weight=0.7, kmeans_loss=Z, kmeans_lambda=0.1, bsize=B)
    model.train()
    out = _with_eps(eps, lambda: model(x))
    if fut:
        pred, future, z, mu, logvar = out
        fl = rv.future_reconstruction_loss(xf, future, "sum")
    else:
        pred, z, mu, logvar = out
        future, fl = None, None
    rec = rv.reconstruction_loss(x, pred, "sum")
    kl = rv.kullback_leibler_loss(mu, logvar)
    km = rv.cluster_loss(z.T, hp["kmeans_loss"], hp["kmeans_lambda"], hp["bsize"])
    total = rec + hp["beta"] * hp["kl_weight"] * kl + hp["kl_weight"] * km
    if fut:
        total = total + fl
    model.zero_grad()
    total.backward()
    d = dict(x=x.numpy(), fut=xf.numpy(), eps=eps.numpy(), pred=pred.detach().numpy(), z=z.detach().numpy(),
             mu=mu.detach().numpy(), logvar=logvar.detach().numpy(),
             loss_rec=rec.item(), loss_kl=kl.item(), loss_kmeans=km.item(), loss_total=total.item(),
             cfg=np.array([B, T, F, Z, H, int(fut), S]), hp_kl_weight=hp["kl_weight"])
    if fut:
        d["future"] = future.detach().numpy()
        d["loss_fut"] = fl.item()
    # eval-mode forward too (z = mu)
    model.eval()
    with torch.no_grad():
        oe = model(x)
    d["pred_eval"] = oe[0].numpy()
    big = H > 64
    for k, p in model.named_parameters():
        g = p.grad.detach().numpy().copy()
        if big:     # 2.6 M params: keep summaries + a strided sample, not 10 MB of gradients
            d["gsum/" + k] = np.array([g.sum(dtype=np.float64), np.abs(g).sum(dtype=np.float64), np.abs(g).max()])
            d["gsample/" + k] = g.reshape(-1)[::97][:512].copy()
        else:
            d["grad/" + k] = g
            d["w/" + k] = p.detach().numpy().copy()
    # three AMSGrad steps on the same batch through the reference objects (rnn_vae.py:332,141-143)
    if not big:
        model.train()
        opt = torch.optim.Adam(model.parameters(), lr=5e-4, amsgrad=True)
        for _ in range(3):
            o = _with_eps(eps, lambda: model(x))
            if fut:
                p_, f_, z_, mu_, lv_ = o
                loss = rv.reconstruction_loss(x, p_, "sum") + rv.future_reconstruction_loss(xf, f_, "sum")
            else:
                p_, z_, mu_, lv_ = o
                loss = rv.reconstruction_loss(x, p_, "sum")
            loss = loss + hp["beta"] * hp["kl_weight"] * rv.kullback_leibler_loss(mu_, lv_) \
                + hp["kl_weight"] * rv.cluster_loss(z_.T, Z, 0.1, B)
            opt.zero_grad()
            loss.backward()
            opt.step()
        for k, p in model.named_parameters():
            d["w3/" + k] = p.detach().numpy()
        d["loss_after3"] = loss.item()
    np.savez_compressed(os.path.join(OUT, "step_%s.npz" % name), **d)
    print("wrote step_%s" % name, "total", total.item())


def gen_step_opts(rm, rv):
    """The option branches the default configuration never takes, through the reference's own objects: Lambda softplus
    (rnn_model.py:59-61,67), MSE reduction 'mean' for both reconstruction terms (rnn_vae.py:35-43), kmeans_loss < zdims
    (rnn_vae.py:48: only the leading singular values) and three different hidden sizes (rnn_model.py:158-161)."""
    B, T, F, Z, S = 20, 6, 7, 6, 3
    H1, HR, HP, K = 64, 32, 96, 3
    torch.manual_seed(19)
    model = rm.RNN_VAE(2 * T, Z, F, True, S, H1, H1, HR, HP, 0, 0, 0, True)
    x, xf, eps = _inputs(B, T, F, S, Z)
    klw, lam = 0.7, 0.1
    model.train()
    pred, future, z, mu, logvar = _with_eps(eps, lambda: model(x))
    rec = rv.reconstruction_loss(x, pred, "mean")
    fl = rv.future_reconstruction_loss(xf, future, "mean")
    kl = rv.kullback_leibler_loss(mu, logvar)
    km = rv.cluster_loss(z.T, K, lam, B)
    total = rec + fl + 1.0 * klw * kl + klw * km
    model.zero_grad()
    total.backward()
    d = dict(x=x.numpy(), fut=xf.numpy(), eps=eps.numpy(), pred=pred.detach().numpy(), future=future.detach().numpy(),
             z=z.detach().numpy(), mu=mu.detach().numpy(), logvar=logvar.detach().numpy(), loss_rec=rec.item(), loss_fut=fl.item(),
             loss_kl=kl.item(), loss_kmeans=km.item(), loss_total=total.item(),
             cfg=np.array([B, T, F, Z, S, H1, HR, HP, K]), hp_kl_weight=klw, hp_lambda=lam)
    model.eval()
    with torch.no_grad():
        d["pred_eval"] = model(x)[0].numpy()
    for k, p in model.named_parameters():
        d["grad/" + k] = p.grad.detach().numpy().copy()
        d["w/" + k] = p.detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, "step_opts.npz"), **d)
    print("wrote step_opts total", total.item())


def gen_train_fn(rm, rv):
    """Reference train()/test() (rnn_vae.py:94-210) on a fixed 5-batch loader; pins the 6-tuple / 3-tuple
    return values including the division by idx = n_batches-1."""
    B, T, F, Z, H, S = 8, 6, 5, 4, 32, 3
    model = _ref_model(rm, T, Z, F, True, S, H)
    g = torch.Generator().manual_seed(7)
    batches = [torch.randn(B, F, 2 * T, generator=g, dtype=torch.float64) for _ in range(5)]
    opt = torch.optim.Adam(model.parameters(), lr=5e-4, amsgrad=True)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=100, gamma=1)
    torch.manual_seed(123)                 # fixes Lambda's randn_like draws
    ret = rv.train(batches, 3, model, opt, "linear", 1, 0, 4, 2 * T, True, S, sched, "sum", "sum", Z, 0.1, B, False)
    torch.manual_seed(321)
    ret_t = rv.test(batches, 3, model, opt, 1, ret[0], 2 * T, "sum", Z, 0.1, True, B)
    d = dict(batches=np.stack([b.numpy() for b in batches]), train_ret=np.array([float(r) for r in ret]),
             test_ret=np.array([float(r) for r in ret_t]), cfg=np.array([B, T, F, Z, H, 1, S]))
    for k, p in model.named_parameters():
        d["w_after/" + k] = p.detach().numpy()
    np.savez_compressed(os.path.join(OUT, "train_fn.npz"), **d)
    print("wrote train_fn", ret)


def gen_embed_synth(rm, ps):
    """Reference embedd_latent_vectors (pose_segmentation.py:67-101) on a synthetic series."""
    T, F, Z, H = 30, 12, 30, 256
    model = _ref_model(rm, T, Z, F, True, 15, H)
    model.eval()
    rng = np.random.default_rng(5)
    N = 400
    series = np.cumsum(rng.standard_normal((F, N)), axis=1) * 0.3      # smooth-ish pose-like series, f64
    series = (series - series.mean()) / series.std()
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "data", "synth"))
        np.save(os.path.join(tmp, "data", "synth", "synth-PE-seq-clean.npy"), series)
        cfg = dict(project_path=tmp, time_window=T, num_features=F)
        lat = ps.embedd_latent_vectors(cfg, ["synth"], model, True)[0]
    np.savez_compressed(os.path.join(OUT, "embed_synth.npz"), series=series, latent=lat,
                        cfg=np.array([T, F, Z, H]))
    print("wrote embed_synth", lat.shape)


def gen_video1(rm, ps):
    """BASELINE config 1 plumbing: examples/video-1.csv -> csv_to_numpy -> create_trainset (reference code,
    egocentric_data=true, F=12) -> embedd_latent_vectors with the seed-19 model.  The cleaned series is
    committed (float32, it is the GPU test's input); the latent vectors are committed for every 40th window."""
    import vame
    from vame.util import auxiliary
    T, F, Z, H = 30, 12, 30, 256
    with tempfile.TemporaryDirectory() as tmp:
        proj = os.path.join(tmp, "proj")
        for sub in ("videos/pose_estimation", "data/video-1", "data/train", "results/video-1", "model"):
            os.makedirs(os.path.join(proj, sub))
        import shutil
        shutil.copy("/root/reference/examples/video-1.csv", os.path.join(proj, "videos", "pose_estimation", "video-1.csv"))
        cfg_file, _ = auxiliary.create_config_template()
        cfg_file["Project"] = "proj"
        cfg_file["project_path"] = proj
        cfg_file["video_sets"] = ["video-1"]
        cfg_file.update(dict(egocentric_data=True, num_features=F, time_window=T, zdims=Z, test_fraction=0.1,
                             pose_confidence=0.99, iqr_factor=4, savgol_filter=True, savgol_length=5,
                             savgol_order=2, robust=True, all_data="yes", batch_size=32,
                             prediction_decoder=1, prediction_steps=15))
        cfgp = os.path.join(proj, "config.yaml")
        auxiliary.write_config(cfgp, cfg_file)
        vame.csv_to_numpy(cfgp)
        vame.create_trainset(cfgp, check_parameter=False)
        clean = np.load(os.path.join(proj, "data", "video-1", "video-1-PE-seq-clean.npy"))
        model = _ref_model(rm, T, Z, F, True, 15, H)
        model.eval()
        cfg = dict(project_path=proj, time_window=T, num_features=F)
        # the reference loop costs ~5 ms/window: embed the first 6000 frames through the literal loop
        part = clean[:, :6000]
        np.save(os.path.join(proj, "data", "video-1", "video-1-PE-seq-clean.npy"), part)
        lat = ps.embedd_latent_vectors(cfg, ["video-1"], model, True)[0]
    np.savez_compressed(os.path.join(OUT, "video1.npz"), clean=clean.astype(np.float32),
                        latent_first=lat[::40].copy(), n_ref_windows=lat.shape[0], cfg=np.array([T, F, Z, H]))
    print("wrote video1", clean.shape, lat.shape)


def gen_video1_trained(rm, rv, ps, epochs=10, batch=128):
    """A TRAINED checkpoint for the video-1 embedding check (the error of split-precision products grows with |W_hh|,
    which the seed-19 initial weights do not exercise): the reference's own train() (vame/model/rnn_vae.py:94-164) on the
    reference's SEQUENCE_DATASET + DataLoader over video-1's train_seq.npy for `epochs` epochs (KL annealing as in the default
    config: kl_start 2, annealtime 4), then the reference's embedd_latent_vectors loop with the trained weights on the first 6000
    frames.  Committed: the state_dict (fp32) and every 40th latent vector; the cleaned series is the one of video1.npz."""
    import shutil
    import torch.utils.data as Data
    import vame
    from vame.util import auxiliary
    T, F, Z, H = 30, 12, 30, 256
    dl = sys.modules["vame.model.dataloader"]
    with tempfile.TemporaryDirectory() as tmp:
        proj = os.path.join(tmp, "proj")
        for sub in ("videos/pose_estimation", "data/video-1", "data/train", "results/video-1", "model"):
            os.makedirs(os.path.join(proj, sub))
        shutil.copy("/root/reference/examples/video-1.csv", os.path.join(proj, "videos", "pose_estimation", "video-1.csv"))
        cfg_file, _ = auxiliary.create_config_template()
        cfg_file["Project"] = "proj"
        cfg_file["project_path"] = proj
        cfg_file["video_sets"] = ["video-1"]
        cfg_file.update(dict(egocentric_data=True, num_features=F, time_window=T, zdims=Z, test_fraction=0.1,
                             pose_confidence=0.99, iqr_factor=4, savgol_filter=True, savgol_length=5,
                             savgol_order=2, robust=True, all_data="yes", batch_size=batch,
                             prediction_decoder=0, prediction_steps=15))
        cfgp = os.path.join(proj, "config.yaml")
        auxiliary.write_config(cfgp, cfg_file)
        vame.csv_to_numpy(cfgp)
        vame.create_trainset(cfgp, check_parameter=False)
        clean = np.load(os.path.join(proj, "data", "video-1", "video-1-PE-seq-clean.npy"))
        model = _ref_model(rm, T, Z, F, False, 0, H)
        np.random.seed(19)                                   # the reference dataset draws window starts from numpy's global stream
        torch.manual_seed(19)
        trainset = dl.SEQUENCE_DATASET(os.path.join(proj, "data", "train", ""), data="train_seq.npy", train=True, temporal_window=2 * T)
        loader = Data.DataLoader(trainset, batch_size=batch, shuffle=True, drop_last=True)
        opt = torch.optim.Adam(model.parameters(), lr=5e-4, amsgrad=True)
        sched = torch.optim.lr_scheduler.StepLR(opt, step_size=100, gamma=1)
        hist = []
        for epoch in range(1, epochs + 1):
            ret = rv.train(loader, epoch, model, opt, "linear", 1, 2, 4, 2 * T, False, 15, sched, "sum", "sum", Z, 0.1, batch, False)
            hist.append([float(v) for v in ret])
            print("epoch", epoch, hist[-1], flush=True)
        model.eval()
        cfg = dict(project_path=proj, time_window=T, num_features=F)
        np.save(os.path.join(proj, "data", "video-1", "video-1-PE-seq-clean.npy"), clean[:, :6000])
        lat = ps.embedd_latent_vectors(cfg, ["video-1"], model, True)[0]
    sd = {"w/" + k: v.detach().numpy() for k, v in model.state_dict().items()}
    whh = max(float(np.abs(v).max()) for k, v in sd.items() if "weight_hh" in k)
    np.savez_compressed(os.path.join(OUT, "video1_trained.npz"), latent_first=lat[::40].copy(), n_ref_windows=lat.shape[0],
                        cfg=np.array([T, F, Z, H]), history=np.array(hist), epochs=np.array([epochs]), batch=np.array([batch]), **sd)
    print("wrote video1_trained", lat.shape, "max |W_hh| %.3f (init bound %.4f)" % (whh, 1 / np.sqrt(H)))


def gen_kmeans(ps):
    """k-means step: outputs of the reference's own same_parameterization / individual_parameterization
    (vame/analysis/pose_segmentation.py:127-196, i.e. sklearn KMeans) on synthetic latent vectors of two 'files', plus
    single-run pieces (sklearn kmeans_plusplus seeding, KMeans(init=array, n_init=1)) that pin the Lloyd restatement."""
    from sklearn.cluster import KMeans, kmeans_plusplus
    rng = np.random.RandomState(7)
    k, Z = 8, 30
    cent = rng.randn(k, Z).astype(np.float32) * 0.35
    X = np.concatenate([cent[i] + (0.4 + 0.15 * i) * rng.randn(375, Z).astype(np.float32) for i in range(k)]).astype(np.float32)
    X = X[rng.permutation(X.shape[0])]
    files = ["a", "b"]
    lat = [X[:1800], X[1800:]]
    cfg = {"random_state_kmeans": 42, "random_state_kmeans: ": 42, "n_init_kmeans": 3}
    lab_s, cen_s, use_s = ps.same_parameterization(cfg, files, lat, k, "kmeans")
    lab_i, cen_i, use_i = ps.individual_parameterization(cfg, files, lat, k)
    seed = 12345
    init, idx = kmeans_plusplus(X, k, random_state=seed)
    km = KMeans(n_clusters=k, init=init, n_init=1).fit(X)
    out = {"X": X, "split": np.array([1800]), "k": np.array([k]),
           "same_labels0": lab_s[0], "same_labels1": lab_s[1], "same_centers": cen_s[0], "same_usage0": use_s[0], "same_usage1": use_s[1],
           "ind_labels0": lab_i[0], "ind_labels1": lab_i[1], "ind_centers0": cen_i[0], "ind_centers1": cen_i[1],
           "pp_seed": np.array([seed]), "pp_init": init, "pp_idx": idx,
           "single_labels": km.labels_, "single_centers": km.cluster_centers_, "single_inertia": np.array([km.inertia_]),
           "single_n_iter": np.array([km.n_iter_])}
    np.savez_compressed(os.path.join(OUT, "kmeans_blobs.npz"), **out)
    print("kmeans_blobs: single inertia %.4f n_iter %d" % (km.inertia_, km.n_iter_))


def gen_trainset():
    """create_trainset step: outputs of the reference's own traindata_fixed / traindata_aligned
    (vame/model/create_training.py:94-264) on small inputs: the first 1500 frames of examples/video-1.csv after the
    reference's csv_to_numpy, and two synthetic two-file projects with injected outliers (so that the IQR cut-off and both
    NaN-interpolation variants are exercised)."""
    import shutil
    import vame
    from vame.util import auxiliary
    from vame.model import create_training as ct
    out = {}

    def project(tmp, files, cfg_over):
        proj = os.path.join(tmp, "proj")
        for sub in ["videos/pose_estimation", "data/train", "model"] + ["data/" + f for f in files] + ["results/" + f for f in files]:
            os.makedirs(os.path.join(proj, sub), exist_ok=True)
        cfg_file, _ = auxiliary.create_config_template()
        cfg_file["Project"] = "proj"
        cfg_file["project_path"] = proj
        cfg_file["video_sets"] = list(files)
        cfg_file.update(cfg_over)
        cfgp = os.path.join(proj, "config.yaml")
        auxiliary.write_config(cfgp, cfg_file)
        return proj, cfgp

    def collect(tag, proj, files, raw):
        out[tag + "/train"] = np.load(os.path.join(proj, "data", "train", "train_seq.npy"))
        out[tag + "/test"] = np.load(os.path.join(proj, "data", "train", "test_seq.npy"))
        for i, f in enumerate(files):
            out["%s/raw%d" % (tag, i)] = raw[i]
            out["%s/clean%d" % (tag, i)] = np.load(os.path.join(proj, "data", f, f + "-PE-seq-clean.npy"))

    base = dict(num_features=12, time_window=30, zdims=30, test_fraction=0.1, pose_confidence=0.99, iqr_factor=4, savgol_filter=True,
                savgol_length=5, savgol_order=2, robust=True, all_data="yes", batch_size=32)
    # ---- video-1 through csv_to_numpy (egocentric_data = True -> traindata_fixed)
    with tempfile.TemporaryDirectory() as tmp:
        proj, cfgp = project(tmp, ["video-1"], dict(base, egocentric_data=True))
        shutil.copy("/root/reference/examples/video-1.csv", os.path.join(proj, "videos", "pose_estimation", "video-1.csv"))
        vame.csv_to_numpy(cfgp)
        f = os.path.join(proj, "data", "video-1", "video-1-PE-seq.npy")
        raw = np.load(f)[:, :1500].copy()
        np.save(f, raw)
        vame.create_trainset(cfgp, check_parameter=False)
        collect("video1_fixed", proj, ["video-1"], [raw])
    # ---- synthetic two-file projects with outliers
    rng = np.random.RandomState(11)

    def synth(F, N):
        t = np.arange(N)
        x = np.stack([np.sin(0.01 * (i + 1) * t + i) * (1 + 0.2 * i) + 0.05 * rng.randn(N) for i in range(F)])
        idx = rng.choice(F * N, size=F * N // 150, replace=False)
        x.reshape(-1)[idx] += rng.choice([-1.0, 1.0], size=idx.size) * rng.uniform(6, 20, size=idx.size)   # spikes beyond 4 x IQR
        return x
    for tag, fixed, sav in (("synth_fixed", True, True), ("synth_fixed_nosav", True, False)):
        with tempfile.TemporaryDirectory() as tmp:
            files = ["a", "b"]
            proj, cfgp = project(tmp, files, dict(base, egocentric_data=fixed, num_features=10, savgol_filter=sav, savgol_length=7, savgol_order=3))
            raw = [synth(10, 700), synth(10, 500)]
            for f, r in zip(files, raw):
                np.save(os.path.join(proj, "data", f, f + "-PE-seq.npy"), r)
            vame.create_trainset(cfgp, check_parameter=False)
            collect(tag, proj, files, raw)
    # aligned path: two anchor rows of constant value (zero variance after alignment) that the reference removes
    with tempfile.TemporaryDirectory() as tmp:
        files = ["a", "b"]
        proj, cfgp = project(tmp, files, dict(base, egocentric_data=False, num_features=10))
        raw = []
        for n in (700, 500):
            r = synth(10, n)
            r[3, :] = 0.25
            r[7, :] = 0.25
            raw.append(r)
        for f, r in zip(files, raw):
            np.save(os.path.join(proj, "data", f, f + "-PE-seq.npy"), r)
        vame.create_trainset(cfgp, check_parameter=False)
        collect("synth_aligned", proj, files, raw)
    np.savez_compressed(os.path.join(OUT, "trainset.npz"), **out)
    print("trainset:", {k: v.shape for k, v in out.items() if k.endswith("/train")})


def main():
    os.makedirs(OUT, exist_ok=True)
    load_reference()
    rm, rv, ps = reference_modules()
    torch.set_num_threads(os.cpu_count())
    if "--only-kmeans" in sys.argv:
        gen_kmeans(ps)
        return
    if "--only-trainset" in sys.argv:
        gen_trainset()
        return
    if "--only-step-opts" in sys.argv:
        gen_step_opts(rm, rv)
        return
    if "--only-video1-trained" in sys.argv:
        gen_video1_trained(rm, rv, ps)
        return
    for name in CASES:
        gen_step_case(name, rm, rv)
    gen_step_opts(rm, rv)
    gen_train_fn(rm, rv)
    gen_embed_synth(rm, ps)
    gen_video1(rm, ps)
    gen_video1_trained(rm, rv, ps)
    gen_kmeans(ps)
    gen_trainset()


if __name__ == "__main__":
    main()
