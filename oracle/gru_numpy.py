"""TEST INFRASTRUCTURE ONLY — numpy restatement of the RNN-VAE hot path.

The reference (LINCellularNeuroscience/VAME) contains no arithmetic of its own on
this path: it calls PyTorch (unpinned in VAME.yaml; README.md:20 says "tested on
PyTorch 1.5"; this container has torch 2.11.0).  This file restates the published
algorithms of those PyTorch ops in plain numpy so that the CUDA kernels can be
checked against something that does not share code with either torch or the
kernels.  dtype follows the inputs (pass float64 arrays for an fp64 "truth").

Weights are passed as a dict keyed by the reference's state_dict names
(SURVEY.md §3.4), e.g. ``encoder.encoder_rnn.weight_hh_l0_reverse``.

Reference call sites restated here:
  * nn.GRU(bidirectional, batch_first)        vame/model/rnn_model.py:34-35,41,91-92,106,125-126,141
  * torch.cat of h_n                           vame/model/rnn_model.py:43
  * Lambda (two Linear, reparam)               vame/model/rnn_model.py:56-57,63-76
  * Decoder / Decoder_Future (.view quirk)     vame/model/rnn_model.py:99-109,133-144
  * RNN_VAE.forward                            vame/model/rnn_model.py:162-179
  * reconstruction / future loss (MSELoss)     vame/model/rnn_vae.py:35-43
  * cluster_loss (Gram + SVD)                  vame/model/rnn_vae.py:45-50
  * kullback_leibler_loss                      vame/model/rnn_vae.py:53-60
  * total loss                                 vame/model/rnn_vae.py:124-129,135-138
  * Adam(amsgrad=True)                         vame/model/rnn_vae.py:332,143
  * embedd_latent_vectors                      vame/analysis/pose_segmentation.py:87-98
"""
import numpy as np


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


# --------------------------------------------------------------------------------------
# GRU (PyTorch gate order r, z, n in the stacked 3H rows)
# --------------------------------------------------------------------------------------
def gru_dir_forward(gi, w_hh, b_hh, h0, reverse):
    """One direction of one GRU layer.

    gi   : (B, T, 3H)  input projection x_t @ W_ih^T + b_ih for every t (already computed)
    w_hh : (3H, H), b_hh : (3H,), h0 : (B, H)
    Returns out (B, T, H) and a dict of per-step saved tensors for BPTT.
    Cell (torch.nn.GRU docs):
        r = s(gi_r + gh_r) ; z = s(gi_z + gh_z) ; n = tanh(gi_n + r*gh_n) ; h' = (1-z)*n + z*h
    """
    B, T, G = gi.shape
    H = G // 3
    out = np.zeros((B, T, H), dtype=gi.dtype)
    sv = dict(r=np.zeros((B, T, H), gi.dtype), z=np.zeros((B, T, H), gi.dtype),
              n=np.zeros((B, T, H), gi.dtype), ghn=np.zeros((B, T, H), gi.dtype),
              hprev=np.zeros((B, T, H), gi.dtype))
    h = h0.astype(gi.dtype)
    steps = range(T - 1, -1, -1) if reverse else range(T)
    for t in steps:
        gh = h @ w_hh.T + b_hh
        r = _sigmoid(gi[:, t, :H] + gh[:, :H])
        z = _sigmoid(gi[:, t, H:2 * H] + gh[:, H:2 * H])
        n = np.tanh(gi[:, t, 2 * H:] + r * gh[:, 2 * H:])
        sv["hprev"][:, t] = h
        h = (1.0 - z) * n + z * h
        sv["r"][:, t], sv["z"][:, t], sv["n"][:, t], sv["ghn"][:, t] = r, z, n, gh[:, 2 * H:]
        out[:, t] = h
    return out, h, sv


def gru_dir_backward(dout, dh_last, sv, w_hh, reverse):
    """BPTT for one direction.

    dout    : (B, T, H) gradient wrt every per-step output (may be zeros)
    dh_last : (B, H)    gradient wrt the final hidden state (h_n)
    Returns dgi (B,T,3H), dgh (B,T,3H), dh0 (B,H).
    """
    B, T, H = dout.shape
    dgi = np.zeros((B, T, 3 * H), dout.dtype)
    dgh = np.zeros((B, T, 3 * H), dout.dtype)
    dh = dh_last.astype(dout.dtype).copy()
    steps = range(T) if reverse else range(T - 1, -1, -1)   # opposite of the forward order
    for t in steps:
        dh = dh + dout[:, t]
        r, z, n, ghn, hp = (sv[k][:, t] for k in ("r", "z", "n", "ghn", "hprev"))
        dn = dh * (1.0 - z)
        dz = dh * (hp - n)
        da_n = dn * (1.0 - n * n)
        da_z = dz * z * (1.0 - z)
        da_r = da_n * ghn * r * (1.0 - r)
        dgi[:, t] = np.concatenate([da_r, da_z, da_n], axis=1)
        dgh[:, t] = np.concatenate([da_r, da_z, da_n * r], axis=1)
        dh = dh * z + dgh[:, t] @ w_hh
    return dgi, dgh, dh


def _gru_names(prefix, layer, reverse):
    sfx = "_l%d%s" % (layer, "_reverse" if reverse else "")
    return (prefix + ".weight_ih" + sfx, prefix + ".weight_hh" + sfx,
            prefix + ".bias_ih" + sfx, prefix + ".bias_hh" + sfx)


def bigru_layer_forward(w, prefix, layer, x, h0f, h0b, const_input=False):
    """Both directions of one layer.  x: (B,T,In) or, if const_input, (B,In) broadcast over T
    (the decoder case, rnn_model.py:169-170: ``z.unsqueeze(2).repeat(1,1,T).permute(0,2,1)``)."""
    res = {}
    for d, rev in enumerate((False, True)):
        wi, wh, bi, bh = (w[k] for k in _gru_names(prefix, layer, rev))
        if const_input:
            gi = np.broadcast_to((x @ wi.T + bi)[:, None, :], (x.shape[0], const_input, wi.shape[0])).copy()
        else:
            gi = x @ wi.T + bi
        out, hn, sv = gru_dir_forward(gi, wh, bh, h0f if d == 0 else h0b, rev)
        res[d] = dict(out=out, hn=hn, sv=sv, gi=gi)
    return res


def encoder_forward(w, x, keep=False):
    """rnn_model.py:40-45 — 2-layer bi-GRU, returns cat(h_n[0..3]) = (B, 4H)."""
    B = x.shape[0]
    H = w["encoder.encoder_rnn.weight_hh_l0"].shape[1]
    z0 = np.zeros((B, H), x.dtype)
    l0 = bigru_layer_forward(w, "encoder.encoder_rnn", 0, x, z0, z0)
    x1 = np.concatenate([l0[0]["out"], l0[1]["out"]], axis=2)
    l1 = bigru_layer_forward(w, "encoder.encoder_rnn", 1, x1, z0, z0)
    hidden = np.concatenate([l0[0]["hn"], l0[1]["hn"], l1[0]["hn"], l1[1]["hn"]], axis=1)
    if keep:
        return hidden, dict(l0=l0, l1=l1, x=x, x1=x1)
    return hidden


def lambda_forward(w, hidden, eps=None, softplus=False):
    """rnn_model.py:63-76.  eps=None -> eval mode (z = mu)."""
    mu = hidden @ w["lmbda.hidden_to_mean.weight"].T + w["lmbda.hidden_to_mean.bias"]
    lv_lin = hidden @ w["lmbda.hidden_to_logvar.weight"].T + w["lmbda.hidden_to_logvar.bias"]
    logvar = np.where(lv_lin > 20, lv_lin, np.log1p(np.exp(np.minimum(lv_lin, 20)))) if softplus else lv_lin
    if eps is None:
        return mu, mu, logvar, lv_lin
    z = eps * np.exp(0.5 * logvar) + mu
    return z, mu, logvar, lv_lin


def decoder_forward(w, name, rnn, z, steps, keep=False):
    """rnn_model.py:99-109 / 133-144.  name in {"decoder","decoder_future"}, rnn in {"rnn_rec","rnn_pred"}.

    h0 quirk: ``latent_to_hidden(z)`` is (B, 2H); ``.view(2, B, H)`` is a raw reshape of the
    row-major buffer (NOT a transpose), so initial states are mixed across batch neighbours."""
    B = z.shape[0]
    hid = z @ w[name + ".latent_to_hidden.weight"].T + w[name + ".latent_to_hidden.bias"]
    H = hid.shape[1] // 2
    h0 = hid.reshape(2, B, H)
    layer = bigru_layer_forward(w, name + "." + rnn, 0, z, h0[0], h0[1], const_input=steps)
    dec_out = np.concatenate([layer[0]["out"], layer[1]["out"]], axis=2)
    pred = dec_out @ w[name + ".hidden_to_output.weight"].T + w[name + ".hidden_to_output.bias"]
    if keep:
        return pred, dict(layer=layer, dec_out=dec_out, hid=hid)
    return pred


def cluster_loss(latent, kloss, lmbda, bsize, return_grad=False):
    """rnn_vae.py:45-50 called as cluster_loss(latent.T, ...) (rnn_vae.py:126,137):
    H = latent.T (Z,B); gram = H.T @ H / bsize (B x B); sv = sqrt(svd(gram)[:k]); lmbda*sum(sv).
    The non-zero singular values of latent@latent.T/bsize equal the eigenvalues of the ZxZ matrix
    latent.T@latent/bsize, so the ZxZ form is used (k is clipped to the rank bound min(B,Z))."""
    B, Z = latent.shape
    G = latent.T @ latent / bsize
    ev, V = np.linalg.eigh(G.astype(np.float64))
    order = np.argsort(ev)[::-1]
    k = min(kloss, B, Z)
    ev_k = np.maximum(ev[order[:k]], 0.0)
    V_k = V[:, order[:k]]
    loss = lmbda * np.sum(np.sqrt(ev_k))
    if not return_grad:
        return loss.astype(latent.dtype) if hasattr(loss, "astype") else loss
    # d/dG sum sqrt(ev_i) = V diag(0.5/sqrt(ev)) V^T ; G = L^T L / bsize -> dL = 2 L S / bsize
    S = (V_k * (0.5 / np.sqrt(ev_k))) @ V_k.T
    dlatent = lmbda * 2.0 * latent.astype(np.float64) @ S / bsize
    return loss, dlatent.astype(latent.dtype)


def cluster_loss_bxb_svd(latent, kloss, lmbda, bsize):
    """Literal form of rnn_vae.py:45-50 (B x B Gram, SVD) for cross-checking the ZxZ form."""
    Hm = latent.T
    gram = (Hm.T @ Hm) / bsize
    sv2 = np.linalg.svd(gram.astype(np.float64), compute_uv=False)
    return lmbda * np.sum(np.sqrt(sv2[:kloss]))


def kl_loss(mu, logvar):
    """rnn_vae.py:53-60."""
    return -0.5 * np.mean(1.0 + logvar - mu ** 2 - np.exp(logvar))


def mse_loss(x, x_tilde, reduction="sum"):
    """rnn_vae.py:35-43 (nn.MSELoss)."""
    d = (x_tilde - x) ** 2
    return d.sum() if reduction == "sum" else d.mean()


# --------------------------------------------------------------------------------------
# Full train-step forward + manual backward
# --------------------------------------------------------------------------------------
def train_step(w, x, fut, eps, hp):
    """Forward + losses + gradients of one batch (rnn_vae.py:106-143 without the optimizer).

    hp: dict(future, steps_future, beta, kl_weight, kmeans_loss, kmeans_lambda, bsize,
             mse_red, mse_pred, softplus)
    Returns (losses dict, grads dict keyed like w, aux dict)."""
    dt = x.dtype
    B, T, F = x.shape
    future = bool(hp.get("future", False))
    S = hp.get("steps_future", 0)
    hidden, enc = encoder_forward(w, x, keep=True)
    z, mu, logvar, lv_lin = lambda_forward(w, hidden, eps, hp.get("softplus", False))
    pred, dec = decoder_forward(w, "decoder", "rnn_rec", z, T, keep=True)
    losses = {}
    losses["rec"] = mse_loss(x, pred, hp.get("mse_red", "sum"))
    if future:
        predf, decf = decoder_forward(w, "decoder_future", "rnn_pred", z, S, keep=True)
        losses["fut"] = mse_loss(fut, predf, hp.get("mse_pred", "sum"))
    losses["kl"] = kl_loss(mu, logvar)
    km, dz_km = cluster_loss(z, hp["kmeans_loss"], hp["kmeans_lambda"], hp["bsize"], return_grad=True)
    losses["kmeans"] = km
    klw, beta = hp["kl_weight"], hp["beta"]
    losses["total"] = losses["rec"] + (losses["fut"] if future else 0.0) + beta * klw * losses["kl"] + klw * km

    g = {}
    dz = klw * dz_km.astype(dt)

    def dec_backward(name, rnn, predx, target, cache, steps, red):
        nonlocal dz
        scale = 2.0 if red == "sum" else 2.0 / predx.size
        dpred = scale * (predx - target)                                   # (B,steps,F)
        Wout = w[name + ".hidden_to_output.weight"]
        g[name + ".hidden_to_output.weight"] = dpred.reshape(-1, F).T @ cache["dec_out"].reshape(-1, Wout.shape[1])
        g[name + ".hidden_to_output.bias"] = dpred.reshape(-1, F).sum(0)
        ddec = dpred @ Wout                                                # (B,steps,2H)
        H = Wout.shape[1] // 2
        dh0 = np.zeros((2, B, H), dt)
        for d, rev in enumerate((False, True)):
            wi, wh, bi, bh = _gru_names(name + "." + rnn, 0, rev)
            lay = cache["layer"][d]
            dgi, dgh, dh0[d] = gru_dir_backward(ddec[:, :, d * H:(d + 1) * H], np.zeros((B, H), dt), lay["sv"], w[wh], rev)
            dgi_sum = dgi.sum(1)                                           # input is the same z at every t
            g[wi] = dgi_sum.T @ z
            g[bi] = dgi.sum((0, 1))
            g[wh] = np.einsum("btg,bth->gh", dgh, lay["sv"]["hprev"])
            g[bh] = dgh.sum((0, 1))
            dz = dz + dgi_sum @ w[wi]
        dhid = dh0.reshape(B, 2 * H)                                       # inverse of the .view quirk
        g[name + ".latent_to_hidden.weight"] = dhid.T @ z
        g[name + ".latent_to_hidden.bias"] = dhid.sum(0)
        dz = dz + dhid @ w[name + ".latent_to_hidden.weight"]

    dec_backward("decoder", "rnn_rec", pred, x, dec, T, hp.get("mse_red", "sum"))
    if future:
        dec_backward("decoder_future", "rnn_pred", predf, fut, decf, S, hp.get("mse_pred", "sum"))

    # Lambda backward
    Z = mu.shape[1]
    c = beta * klw / (B * Z)
    dmu = dz + c * mu
    dlogvar = dz * eps * 0.5 * np.exp(0.5 * logvar) + c * 0.5 * (np.exp(logvar) - 1.0)
    dlv_lin = dlogvar * _sigmoid(lv_lin) if hp.get("softplus", False) else dlogvar
    g["lmbda.hidden_to_mean.weight"] = dmu.T @ hidden
    g["lmbda.hidden_to_mean.bias"] = dmu.sum(0)
    g["lmbda.hidden_to_logvar.weight"] = dlv_lin.T @ hidden
    g["lmbda.hidden_to_logvar.bias"] = dlv_lin.sum(0)
    dhidden = dmu @ w["lmbda.hidden_to_mean.weight"] + dlv_lin @ w["lmbda.hidden_to_logvar.weight"]

    # Encoder backward: only h_n is used (rnn_model.py:41-43), per-step outputs of layer 1 get no gradient
    H = hidden.shape[1] // 4
    zeros_o = np.zeros((B, T, H), dt)
    dx1 = np.zeros((B, T, 2 * H), dt)
    for d, rev in enumerate((False, True)):
        wi, wh, bi, bh = _gru_names("encoder.encoder_rnn", 1, rev)
        lay = enc["l1"][d]
        dgi, dgh, _ = gru_dir_backward(zeros_o, dhidden[:, (2 + d) * H:(3 + d) * H], lay["sv"], w[wh], rev)
        g[wi] = np.einsum("btg,bti->gi", dgi, enc["x1"])
        g[bi] = dgi.sum((0, 1))
        g[wh] = np.einsum("btg,bth->gh", dgh, lay["sv"]["hprev"])
        g[bh] = dgh.sum((0, 1))
        dx1 = dx1 + dgi @ w[wi]
    for d, rev in enumerate((False, True)):
        wi, wh, bi, bh = _gru_names("encoder.encoder_rnn", 0, rev)
        lay = enc["l0"][d]
        dgi, dgh, _ = gru_dir_backward(dx1[:, :, d * H:(d + 1) * H], dhidden[:, d * H:(d + 1) * H], lay["sv"], w[wh], rev)
        g[wi] = np.einsum("btg,bti->gi", dgi, x)
        g[bi] = dgi.sum((0, 1))
        g[wh] = np.einsum("btg,bth->gh", dgh, lay["sv"]["hprev"])
        g[bh] = dgh.sum((0, 1))
    aux = dict(pred=pred, z=z, mu=mu, logvar=logvar, hidden=hidden)
    if future:
        aux["future"] = predf
    return losses, g, aux


def amsgrad_step(p, g, m, v, vmax, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam(amsgrad=True, weight_decay=0) single-tensor update (rnn_vae.py:332,143).
    ``step`` is the 1-based step count AFTER increment.  Updates in place and returns p."""
    m *= beta1
    m += (1 - beta1) * g
    v *= beta2
    v += (1 - beta2) * g * g
    np.maximum(vmax, v, out=vmax)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = np.sqrt(vmax) / np.sqrt(bc2) + eps
    p -= (lr / bc1) * (m / denom)
    return p


def embed_windows(w, series, T, batch=4096):
    """pose_segmentation.py:87-98: for i in range(N-T): mu = lmbda(encoder(series[:, i:i+T].T)).
    series: (F, N).  Returns (N-T, Z).  Windows are independent (h0 = 0), so they are batched."""
    F, N = series.shape
    outs = []
    for s in range(0, N - T, batch):
        e = min(N - T, s + batch)
        idx = np.arange(s, e)[:, None] + np.arange(T)[None, :]
        xw = series.T[idx]                                     # (b, T, F)
        hidden = encoder_forward(w, xw.astype(w["lmbda.hidden_to_mean.weight"].dtype))
        outs.append(lambda_forward(w, hidden, None)[1])
    return np.concatenate(outs, axis=0)
