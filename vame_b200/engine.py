"""Host-side engine: owns the flat parameter / gradient / optimizer buffers and the workspaces, and drives the
C-ABI (include/vame_b200.h) through ctypes.  PyTorch is used for device memory, streams and torch.distributed only;
every FLOP of the hot path runs in libvame_b200.so.  There is no CPU fallback.
"""
import ctypes
import os
from collections import OrderedDict

import torch

from . import _lib as L


class VameDims(ctypes.Structure):
    _fields_ = [("num_features", ctypes.c_int), ("time_window", ctypes.c_int), ("zdims", ctypes.c_int),
                ("hidden_enc", ctypes.c_int), ("hidden_rec", ctypes.c_int), ("hidden_pred", ctypes.c_int),
                ("future_decoder", ctypes.c_int), ("future_steps", ctypes.c_int), ("softplus", ctypes.c_int)]


class VameLossCfg(ctypes.Structure):
    _fields_ = [("mse_red_mean", ctypes.c_int), ("mse_pred_mean", ctypes.c_int), ("kmeans_loss", ctypes.c_int),
                ("kmeans_lambda", ctypes.c_float), ("bsize", ctypes.c_float), ("beta", ctypes.c_float),
                ("kl_weight", ctypes.c_float), ("with_future", ctypes.c_int), ("defer_prior_join", ctypes.c_int)]


HY_LR, HY_KLW, HY_BETA, HY_KMLAMBDA = 0, 1, 2, 3


def state_dict_names(future_decoder):
    """Parameter names in the reference's state_dict order (vame/model/rnn_model.py)."""
    def gru(prefix, layers):
        out = []
        for l in range(layers):
            for sfx in ("", "_reverse"):
                for p in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
                    out.append("%s.%s_l%d%s" % (prefix, p, l, sfx))
        return out
    names = gru("encoder.encoder_rnn", 2)
    names += ["lmbda.hidden_to_mean.weight", "lmbda.hidden_to_mean.bias", "lmbda.hidden_to_logvar.weight", "lmbda.hidden_to_logvar.bias"]
    names += gru("decoder.rnn_rec", 1) + ["decoder.latent_to_hidden.weight", "decoder.latent_to_hidden.bias",
                                          "decoder.hidden_to_output.weight", "decoder.hidden_to_output.bias"]
    if future_decoder:
        names += gru("decoder_future.rnn_pred", 1) + ["decoder_future.latent_to_hidden.weight", "decoder_future.latent_to_hidden.bias",
                                                      "decoder_future.hidden_to_output.weight", "decoder_future.hidden_to_output.bias"]
    return names


def param_shapes(d):
    F, Z, H, Hr, Hp = d.num_features, d.zdims, d.hidden_enc, d.hidden_rec, d.hidden_pred

    def gru(In, Hh, layers):
        out = []
        for l in range(layers):
            i = In if l == 0 else 2 * Hh
            out += [(3 * Hh, i), (3 * Hh, Hh), (3 * Hh,), (3 * Hh,)] * 2
        return out
    shapes = gru(F, H, 2) + [(Z, 4 * H), (Z,), (Z, 4 * H), (Z,)]
    shapes += gru(Z, Hr, 1) + [(2 * Hr, Z), (2 * Hr,), (F, 2 * Hr), (F,)]
    if d.future_decoder:
        shapes += gru(Z, Hp, 1) + [(2 * Hp, Z), (2 * Hp,), (F, 2 * Hp), (F,)]
    return shapes


class Engine:
    """Flat-buffer owner + C-ABI driver for one RNN-VAE instance on one CUDA device."""

    def __init__(self, num_features, time_window, zdims, hidden_enc=256, hidden_rec=256, hidden_pred=256,
                 future_decoder=False, future_steps=0, softplus=False, device=None):
        self.lib = L.lib()
        self.dims = VameDims(int(num_features), int(time_window), int(zdims), int(hidden_enc), int(hidden_rec), int(hidden_pred),
                             int(bool(future_decoder)), int(future_steps or 0), int(bool(softplus)))
        n = self.lib.vame_param_tensors(ctypes.byref(self.dims))
        offs = (ctypes.c_long * n)()
        sizes = (ctypes.c_long * n)()
        total = self.lib.vame_param_layout(ctypes.byref(self.dims), offs, sizes)
        if total <= 0:
            raise L.VameB200Error("vame_param_layout: %s" % self.lib.vame_last_error().decode())
        self.names = state_dict_names(bool(future_decoder))
        self.shapes = param_shapes(self.dims)
        assert len(self.names) == n == len(self.shapes)
        self.offsets = [int(o) for o in offs]
        self.sizes = [int(s) for s in sizes]
        for s, shp in zip(self.sizes, self.shapes):
            numel = 1
            for v in shp:
                numel *= v
            assert numel == s, (shp, s)
        self.n_flat = int(total)
        self.device = torch.device(device) if device is not None else None
        self.flat = None
        self.grad = None
        self.packed = None
        self._packed_version = None
        self._external_dirty = True        # weights changed outside TrainStep since the last FULL re-pack
        self._watched = ()                 # tensors whose version counters are part of the weight version (nn.Parameters bound to flat)
        self._train_gen = {}               # batch size -> generation of the saved activations in the training workspace
        self._ws = {}
        self.hyper = None
        self.opt_state = None
        if self.device is not None and self.device.type == "cuda":
            self.allocate(self.device)

    # ---- memory ---------------------------------------------------------------------------------------
    def allocate(self, device):
        if not torch.cuda.is_available():
            raise L.VameB200Error("vame_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device = torch.device(device)
        self.flat = torch.zeros(self.n_flat, dtype=torch.float32, device=self.device)
        self.grad = torch.zeros(self.n_flat, dtype=torch.float32, device=self.device)
        nb = self.lib.vame_packed_weights_bytes(ctypes.byref(self.dims))
        self.packed = torch.zeros(nb, dtype=torch.uint8, device=self.device)
        self.hyper = torch.zeros(8, dtype=torch.float32, device=self.device)
        self.losses = torch.zeros(8, dtype=torch.float32, device=self.device)
        self._packed_version = None
        self._ws = {}

    def views(self, flat=None):
        flat = self.flat if flat is None else flat
        return OrderedDict((n, flat[o:o + s].view(shp)) for n, o, s, shp in zip(self.names, self.offsets, self.sizes, self.shapes))

    def load_state_dict(self, sd):
        v = self.views()
        with torch.no_grad():
            for k in self.names:
                v[k].copy_(torch.as_tensor(sd[k]).to(self.device, torch.float32))
        self.mark_dirty()

    def state_dict(self):
        return OrderedDict((k, t.detach().clone()) for k, t in self.views().items())

    def _dev(self):
        """Every C-ABI call is enqueued on the current stream of the ENGINE's device, whatever the caller's current device is."""
        return torch.cuda.device(self.device)

    def watch(self, tensors):
        """Tensors that alias ``flat`` but keep their own version counters (``p.data = view`` gives the nn.Parameter its own
        counter: optimizer.step() / p.copy_() bump p._version, not flat._version).  Their counters become part of the version
        that decides whether the packed tensor-core copies are stale."""
        self._watched = tuple(tensors)

    def _version(self):
        return (self.flat._version,) + tuple(t._version for t in self._watched)

    def mark_dirty(self):
        """The fp32 weights changed behind the engine's back (load_state_dict, manual edits): full re-pack on next use."""
        self._packed_version = None
        self._external_dirty = True

    def _ensure_packed(self, force=False):
        ver = self._version()
        if force or self._packed_version != ver:
            self.pack_weights()

    def pack_weights(self):
        """Unconditionally refresh the tensor-core weight copies on the current stream (graph-capturable)."""
        with self._dev():
            L.check(self.lib.vame_pack_weights(ctypes.byref(self.dims), L.ptr(self.flat), L.ptr(self.packed), L.cur_stream()),
                    "vame_pack_weights")
        self._packed_version = self._version()
        self._external_dirty = False

    def workspace(self, batch, training):
        key = (int(batch), bool(training))
        ws = self._ws.get(key)
        if ws is None:
            nb = self.lib.vame_workspace_bytes(ctypes.byref(self.dims), int(batch), int(training))
            if nb == 0:
                raise L.VameB200Error("vame_workspace_bytes failed: %s" % self.lib.vame_last_error().decode())
            # zero-filled: the library keeps small pieces of state in a training workspace between calls (the eigenvector basis
            # that warm-starts the k-means-prior solver) and recognises "no state yet" deterministically
            ws = torch.zeros(nb, dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws

    # ---- hot path ---------------------------------------------------------------------------------------
    def forward(self, x, eps=None, save=False, want=("pred", "future", "z", "mu", "logvar"), ensure_packed=True):
        """RNN_VAE.forward.  x: (B, T, F) float32 CUDA tensor (feature stride 1).  Returns dict of outputs."""
        d = self.dims
        assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 3 and x.shape[1] == d.time_window and x.shape[2] == d.num_features
        if x.stride(2) != 1:
            x = x.contiguous()
        B = x.shape[0]
        if ensure_packed:
            self._ensure_packed()
        ws = self.workspace(B, save)
        out = {}
        if "pred" in want:
            out["pred"] = torch.empty(B, d.time_window, d.num_features, device=self.device)
        if "future" in want and d.future_decoder:
            out["future"] = torch.empty(B, d.future_steps, d.num_features, device=self.device)
        for k in ("z", "mu", "logvar"):
            if k in want:
                out[k] = torch.empty(B, d.zdims, device=self.device)
        if eps is not None:
            eps = eps.contiguous()
        with self._dev():
            L.check(self.lib.vame_forward(ctypes.byref(d), B, L.ptr(self.flat), L.ptr(self.packed), L.ptr(x), x.stride(0), x.stride(1),
                                          L.ptr(eps), int(save), L.ptr(out.get("pred")), L.ptr(out.get("future")), L.ptr(out.get("z")),
                                          L.ptr(out.get("mu")), L.ptr(out.get("logvar")), L.ptr(ws), ws.numel(), L.cur_stream()),
                    "vame_forward")
        self._last = (B, bool(save))
        if save:                             # the training workspace of this batch size now holds THIS forward's activations
            self._train_gen[B] = self._train_gen.get(B, 0) + 1
        return out

    def loss_cfg(self, mse_red="sum", mse_pred="sum", kmeans_loss=None, kmeans_lambda=0.1, bsize=None, beta=1.0, kl_weight=1.0,
                 with_future=True):
        d = self.dims
        return VameLossCfg(int(mse_red == "mean"), int(mse_pred == "mean"), int(d.zdims if kmeans_loss is None else kmeans_loss),
                           float(kmeans_lambda), float(bsize if bsize is not None else 0), float(beta), float(kl_weight),
                           int(bool(with_future) and bool(d.future_decoder)), 0)

    def loss(self, cfg, fut=None, want_grads=True, use_hyper=False, out=None, target=None):
        """Loss terms of the last forward -> device float tensor [rec, fut, kl, kmeans, total, ...]."""
        B, save = self._last
        assert (not want_grads) or save, "loss gradients need forward(save=True)"
        if cfg.bsize == 0:
            cfg.bsize = float(B)
        ws = self.workspace(B, save)
        out = self.losses if out is None else out
        fs0 = fs1 = 0
        if fut is not None:
            if fut.stride(2) != 1:
                fut = fut.contiguous()
            fs0, fs1 = fut.stride(0), fut.stride(1)
        ts0 = ts1 = 0
        if target is not None:
            if target.stride(2) != 1:
                target = target.contiguous()
            ts0, ts1 = target.stride(0), target.stride(1)
        with self._dev():
            L.check(self.lib.vame_loss(ctypes.byref(self.dims), B, ctypes.byref(cfg), L.ptr(fut), fs0, fs1, L.ptr(target), ts0, ts1,
                                       L.ptr(self.hyper) if use_hyper else None, L.ptr(out), int(want_grads), L.ptr(ws), ws.numel(),
                                       L.cur_stream()), "vame_loss")
        return out

    def backward(self, cfg=None, use_loss_grads=True, use_hyper=False, dpred=None, dfuture=None, dz=None, dmu=None, dlogvar=None,
                 batch=None):
        """loss.backward(): fills self.grad (flat, overwritten).  ``batch``: batch size of the forward(save=True) whose saved
        activations are to be used (default: the last forward, which must have been a saving one)."""
        if batch is None:
            B, save = self._last
            assert save, "backward needs forward(save=True)"
        else:
            B = int(batch)
            assert B in self._train_gen, "backward needs forward(save=True) of this batch size"
        ws = self.workspace(B, True)

        def c(t):
            return None if t is None else t.contiguous()
        dpred, dfuture, dz, dmu, dlogvar = c(dpred), c(dfuture), c(dz), c(dmu), c(dlogvar)
        with self._dev():
            L.check(self.lib.vame_backward(ctypes.byref(self.dims), B, L.ptr(self.flat), L.ptr(self.packed), int(use_loss_grads),
                                           ctypes.byref(cfg) if cfg is not None else None, L.ptr(self.hyper) if use_hyper else None,
                                           L.ptr(dpred), L.ptr(dfuture), L.ptr(dz), L.ptr(dmu), L.ptr(dlogvar), L.ptr(self.grad), L.ptr(ws),
                                           ws.numel(), L.cur_stream()), "vame_backward")
        return self.grad

    def init_optimizer(self):
        if self.opt_state is None:
            z = lambda: torch.zeros(self.n_flat, dtype=torch.float32, device=self.device)  # noqa: E731
            self.opt_state = dict(exp_avg=z(), exp_avg_sq=z(), max_exp_avg_sq=z(),
                                  step=torch.zeros(1, dtype=torch.int32, device=self.device),
                                  scratch=torch.zeros(2, dtype=torch.float32, device=self.device))
        return self.opt_state

    def adam_step(self, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0, use_hyper=False, repack=True):
        """torch.optim.Adam(amsgrad=True).step() on the flat buffers (+ refresh of the packed weights)."""
        s = self.init_optimizer()
        with self._dev():
            L.check(self.lib.vame_adam_step(L.ptr(self.flat), L.ptr(self.grad), L.ptr(s["exp_avg"]), L.ptr(s["exp_avg_sq"]),
                                            L.ptr(s["max_exp_avg_sq"]), self.n_flat, float(lr), L.ptr(self.hyper) if use_hyper else None,
                                            L.ptr(s["step"]), L.ptr(s["scratch"]), float(betas[0]), float(betas[1]), float(eps),
                                            float(grad_scale), L.cur_stream()), "vame_adam_step")
        if repack:
            self.pack_weights()
        else:
            self._packed_version = None      # (the caller re-packs what it needs; any other consumer re-packs fully)

    def set_hyper(self, lr=None, kl_weight=None, beta=None, kmeans_lambda=None):
        vals = self.hyper.tolist() if any(v is None for v in (lr, kl_weight, beta, kmeans_lambda)) else [0.0] * 8
        for i, v in ((HY_LR, lr), (HY_KLW, kl_weight), (HY_BETA, beta), (HY_KMLAMBDA, kmeans_lambda)):
            if v is not None:
                vals[i] = float(v)
        self.hyper.copy_(torch.tensor(vals, dtype=torch.float32))

    # ---- inference helpers -------------------------------------------------------------------------------
    def encoder_forward(self, x):
        d = self.dims
        if x.stride(2) != 1:
            x = x.contiguous()
        B = x.shape[0]
        self._ensure_packed()
        ws = self.workspace(B, False)
        hidden = torch.empty(B, 4 * d.hidden_enc, device=self.device)
        with self._dev():
            L.check(self.lib.vame_encoder_forward(ctypes.byref(d), B, L.ptr(self.flat), L.ptr(self.packed), L.ptr(x), x.stride(0), x.stride(1),
                                                  L.ptr(hidden), L.ptr(ws), ws.numel(), L.cur_stream()), "vame_encoder_forward")
        return hidden

    def lambda_forward(self, hidden, eps=None):
        d = self.dims
        hidden = hidden.contiguous()
        B = hidden.shape[0]
        self._ensure_packed()
        ws = self.workspace(B, False)
        z, mu, lv = (torch.empty(B, d.zdims, device=self.device) for _ in range(3))
        with self._dev():
            L.check(self.lib.vame_lambda_forward(ctypes.byref(d), B, L.ptr(self.flat), L.ptr(self.packed), L.ptr(hidden),
                                                 L.ptr(eps.contiguous()) if eps is not None else None, L.ptr(z), L.ptr(mu), L.ptr(lv),
                                                 L.ptr(ws), ws.numel(), L.cur_stream()), "vame_lambda_forward")
        return z, mu, lv

    def decoder_forward(self, z, which=0):
        d = self.dims
        z = z.contiguous()
        B = z.shape[0]
        self._ensure_packed()
        ws = self.workspace(B, False)
        steps = d.time_window if which == 0 else d.future_steps
        pred = torch.empty(B, steps, d.num_features, device=self.device)
        with self._dev():
            L.check(self.lib.vame_decoder_forward(ctypes.byref(d), B, int(which), L.ptr(self.flat), L.ptr(self.packed), L.ptr(z), L.ptr(pred),
                                                  L.ptr(ws), ws.numel(), L.cur_stream()), "vame_decoder_forward")
        return pred

    def embed(self, series_nf, first_window=0, n_windows=None, chunk=9472, out=None):
        """mu of every stride-1 window of a frame-major (N, F) float32 CUDA series -> (n_windows, Z).
        chunk: windows per pass; the default 9472 = 74 tiles of 128 rows x 2 directions = 148 CTAs of the row-resident sweep
        kernel, one per SM."""
        d = self.dims
        assert series_nf.is_cuda and series_nf.dtype == torch.float32 and series_nf.is_contiguous() and series_nf.shape[1] == d.num_features
        N = series_nf.shape[0]
        if n_windows is None:
            n_windows = N - d.time_window - first_window          # range(N - T): pose_segmentation.py:87
        n_windows = max(0, int(n_windows))
        chunk = int(min(chunk, max(128, (n_windows + 127) // 128 * 128)))
        self._ensure_packed()
        key = ("embed", N, chunk)
        ws = self._ws.get(key)
        if ws is None:
            nb = self.lib.vame_embed_workspace_bytes(ctypes.byref(d), N, chunk)
            ws = torch.empty(nb, dtype=torch.uint8, device=self.device)
            self._ws = {k: v for k, v in self._ws.items() if not (isinstance(k[0], str) and k[0] == "embed")}
            self._ws[key] = ws
        if out is None:
            out = torch.empty(n_windows, d.zdims, device=self.device)
        if n_windows > 0:
          with self._dev():
              L.check(self.lib.vame_embed_windows(ctypes.byref(d), L.ptr(self.flat), L.ptr(self.packed), L.ptr(series_nf), N, int(first_window),
                                                    n_windows, chunk, L.ptr(out), L.ptr(ws), ws.numel(), L.cur_stream()), "vame_embed_windows")
        return out

    def cluster_loss(self, latent, kloss, lmbda, bsize, grad_coef=1.0, want_grad=False):
        latent = latent.contiguous()
        B, Z = latent.shape
        loss = torch.zeros(1, dtype=torch.float64, device=latent.device)
        dl = torch.empty_like(latent) if want_grad else None
        with self._dev():
            L.check(self.lib.vame_cluster_loss(L.ptr(latent), B, Z, int(kloss), float(lmbda), float(bsize), float(grad_coef), L.ptr(loss),
                                               L.ptr(dl), L.cur_stream()), "vame_cluster_loss")
        return loss, dl


class TrainStep:
    """One fused data-parallel train step on static buffers:
         forward + losses + backward   ->  [one NCCL sum-allreduce of the flat gradient]  ->  AMSGrad + weight re-pack
    The two compute phases are captured once as CUDA graphs and replayed (the launch-bound recurrence is ~200 dependent
    kernels); hyper-parameters that change between steps (lr, kl_weight, ...) live in device memory (Engine.hyper)."""

    def __init__(self, eng, batch, cfg, world=1, use_graph=True, betas=(0.9, 0.999), eps=1e-8, sampler=None, early_opt=None):
        self.eng, self.B, self.cfg, self.world = eng, int(batch), cfg, int(world)
        # optional vame_b200.dataloader.DeviceWindowSampler: its vame_sample_windows launch becomes the first node of the step
        # (window starts, z-score, data / future split and the reparameterisation noise are produced on the device)
        self.sampler = sampler
        d = eng.dims
        dev = eng.device
        self.x = torch.zeros(batch, d.time_window, d.num_features, device=dev)
        self.fut = torch.zeros(batch, max(d.future_steps, 1), d.num_features, device=dev) if d.future_decoder else None
        self.eps = torch.zeros(batch, d.zdims, device=dev)
        self.losses = torch.zeros(8, device=dev)
        self.acc = torch.zeros(8, dtype=torch.float64, device=dev)   # epoch sums of the loss terms (accumulated inside the step)
        self.betas, self.adam_eps = betas, eps
        self.graphs = None
        self.nccl_in_graph = False
        # N > 1: the allreduce of everything but encoder layer 0's gradients runs on a communication stream under the last BPTT
        # sweep (vame_grad_overlap); VAME_B200_GRAD_OVERLAP=0 falls back to one allreduce of the whole buffer after the backward
        # Measured at N = 2 (profiles/r2_dp_n2.md): +3.2 % at 512 windows per GPU, neutral to slightly negative at 256 (the NCCL
        # CTAs compete with the 128-SM sweep for the 20 SMs the weight-gradient GEMMs run on) -> on by default from 384 up
        ov = os.environ.get("VAME_B200_GRAD_OVERLAP")
        self.overlap = self.world > 1 and (ov != "0" if ov is not None else self.B >= 384)
        # N = 1, opt-in (VAME_B200_EARLY_OPT=1): the optimizer step + weight re-pack of everything but encoder layer 0 on a side stream
        # under the last BPTT sweep (their gradients are final ~200 us before the backward pass ends).  Measured neutral (C2 292.7 k
        # vs 295.7 k, C5 363.6 k vs 362.1 k windows/s): the end of the step is bound by the weight-gradient GEMMs, not by AMSGrad
        self.early_opt = (self.world == 1 and os.environ.get("VAME_B200_EARLY_OPT", "0") != "0"
                          and os.environ.get("VAME_B200_DEFER_REPACK", "0") == "0") if early_opt is None else bool(early_opt)
        self._opt_stream = torch.cuda.Stream(device=dev) if self.early_opt else None
        if self.early_opt:
            self.split = int(eng.lib.vame_grad_bucket_split(ctypes.byref(eng.dims)))
        self._comm_stream = None
        if self.overlap:
            self.split = int(eng.lib.vame_grad_bucket_split(ctypes.byref(eng.dims)))
            self._comm_stream = torch.cuda.Stream(device=dev)
        self.use_graph = use_graph
        self._stage, self._staged, self._k_load = None, [], 0
        self.defer_repack = os.environ.get("VAME_B200_DEFER_REPACK", "0") != "0"
        self.early_prior = os.environ.get("VAME_B200_EARLY_PRIOR", "1") != "0"
        eng.init_optimizer()
        if cfg.bsize == 0:
            cfg.bsize = float(batch)

    def _adam_range(self, lo, hi, stream_ptr):
        e, st = self.eng, self.eng.opt_state
        off = lo * 4
        p = lambda t: ctypes.c_void_p(t.data_ptr() + off)  # noqa: E731
        L.check(e.lib.vame_adam_apply(p(e.flat), p(e.grad), p(st["exp_avg"]), p(st["exp_avg_sq"]), p(st["max_exp_avg_sq"]), hi - lo,
                                      L.ptr(st["scratch"]), float(self.betas[0]), float(self.betas[1]), float(self.adam_eps),
                                      1.0 / self.world, stream_ptr), "vame_adam_apply")

    def _phase1(self):
        e = self.eng
        if self.early_opt:
            # step counter + bias-corrected step size for this step, then fork the optimizer stream (it starts working when the
            # backward pass signals that every gradient outside encoder layer 0 is final)
            st = e.opt_state
            L.check(e.lib.vame_adam_prepare(0.0, L.ptr(e.hyper), L.ptr(st["step"]), L.ptr(st["scratch"]), float(self.betas[0]),
                                            float(self.betas[1]), L.cur_stream()), "vame_adam_prepare")
            self._opt_stream.wait_stream(torch.cuda.current_stream())
        if self.sampler is not None:
            self.sampler.fill(self.x, self.fut if self.cfg.with_future else None, self.eps)
        if self.defer_repack:
            # weights updated by the previous step's optimizer kernel: re-pack what the first sweep needs now, the rest beside it
            L.check(e.lib.vame_pack_weights_deferred(ctypes.byref(e.dims), L.ptr(e.flat), L.ptr(e.packed), L.cur_stream()),
                    "vame_pack_weights_deferred")
        # the k-means prior starts as soon as z exists (inside vame_forward), overlaps the decoders, the losses and the decoder
        # BPTT, and is joined by vame_backward right before the Lambda backward
        self.cfg.defer_prior_join = 1
        if self.cfg.bsize == 0:
            self.cfg.bsize = float(self.B)
        if self.early_prior:
            L.check(e.lib.vame_arm_prior(ctypes.byref(self.cfg), L.ptr(e.hyper)), "vame_arm_prior")
        try:
            e.forward(self.x, self.eps, save=True, want=(), ensure_packed=False)
        finally:
            if self.early_prior:
                e.lib.vame_arm_prior(None, None)
        e.loss(self.cfg, self.fut if self.cfg.with_future else None, want_grads=True, use_hyper=True, out=self.losses)
        if self.overlap or self.early_opt:
            L.check(e.lib.vame_grad_overlap(2 if self.early_opt else 1), "vame_grad_overlap")
        try:
            e.backward(self.cfg, use_hyper=True)
            if self.early_opt:
                so = self._opt_stream
                with torch.cuda.stream(so):
                    sp = ctypes.c_void_p(so.cuda_stream)
                    L.check(e.lib.vame_wait_grads_ready(sp), "vame_wait_grads_ready")
                    self._adam_range(self.split, e.n_flat, sp)
                    L.check(e.lib.vame_pack_weights_train_part(ctypes.byref(e.dims), L.ptr(e.flat), L.ptr(e.packed), self.B, 1, sp),
                            "vame_pack_weights_train_part")
        finally:
            if self.overlap or self.early_opt:
                e.lib.vame_grad_overlap(0)
        self.cfg.defer_prior_join = 0

    def _allreduce(self):
        """Gradient exchange between the two phases (world > 1)."""
        import torch.distributed as dist
        g = self.eng.grad
        if not self.overlap:
            dist.all_reduce(g, op=dist.ReduceOp.SUM)
            return
        cur = torch.cuda.current_stream()
        cs = self._comm_stream
        with torch.cuda.stream(cs):
            # bucket A = everything but encoder layer 0: final when the event inside the (already enqueued) backward fires
            L.check(self.eng.lib.vame_wait_grads_ready(ctypes.c_void_p(cs.cuda_stream)), "vame_wait_grads_ready")
            dist.all_reduce(g[self.split:], op=dist.ReduceOp.SUM)
        dist.all_reduce(g[:self.split], op=dist.ReduceOp.SUM)      # bucket B: after the backward (stream order)
        cur.wait_stream(cs)

    def _phase2(self):
        # with defer_repack the packed copies are refreshed at the start of the next step (and by _ensure_packed for any other caller)
        e = self.eng
        if self.early_opt:
            # encoder layer 0 (its gradients were final last), then join the optimizer stream
            self._adam_range(0, self.split, L.cur_stream())
            L.check(e.lib.vame_pack_weights_train_part(ctypes.byref(e.dims), L.ptr(e.flat), L.ptr(e.packed), self.B, 0, L.cur_stream()),
                    "vame_pack_weights_train_part")
            torch.cuda.current_stream().wait_stream(self._opt_stream)
            e._packed_version = None
            self.acc.add_(self.losses)
            return
        e.adam_step(betas=self.betas, eps=self.adam_eps, grad_scale=1.0 / self.world, use_hyper=True, repack=False)
        if not self.defer_repack:
            # only the weight formats a step of THIS batch size reads; any other consumer re-packs everything (mark_dirty in run())
            L.check(e.lib.vame_pack_weights_train(ctypes.byref(e.dims), L.ptr(e.flat), L.ptr(e.packed), self.B, L.cur_stream()),
                    "vame_pack_weights_train")
        self.acc.add_(self.losses)           # device-side epoch accumulation: no host work per batch

    def capture(self):
        """Warm up eagerly (also refreshes the packed weights), then capture both phases."""
        e = self.eng
        e.pack_weights()
        if not self.use_graph:
            return False
        state = {k: v.clone() for k, v in e.opt_state.items()}
        flat = e.flat.clone()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        try:
            with torch.cuda.stream(s):
                for _ in range(2):
                    self._phase1()
                    if self.world > 1:               # warms up the communicator before any capture
                        self._allreduce()
                    self._phase2()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            if self.world == 1:
                # single GPU: nothing happens between the phases -> ONE graph (one launch instead of two per step)
                g1 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g1):
                    self._phase1()
                    self._phase2()
                self.graphs = (g1, None)
            else:
                # N > 1: the NCCL allreduce is captured INSIDE the graph (one launch per step, no host round trip between
                # backward, allreduce and optimizer); if this NCCL / torch build cannot capture collectives, fall back to two
                # graphs around an eager allreduce
                import torch.distributed as dist
                self.graphs = None
                # (opt-in: measured no faster than two graphs at N = 2, and torch's destroy_process_group() then hangs at exit
                #  waiting for the captured collectives - profiles/r2_dp_n2.md)
                if os.environ.get("VAME_B200_NCCL_IN_GRAPH", "0") != "0":
                    try:
                        g1 = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g1):
                            self._phase1()
                            dist.all_reduce(e.grad, op=dist.ReduceOp.SUM)
                            self._phase2()
                        self.graphs = (g1, None)
                        self.nccl_in_graph = True
                    except Exception as ex:
                        self.capture_error = "nccl-in-graph: " + repr(ex)
                        torch.cuda.synchronize()
                if self.graphs is None:
                    g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g1):
                        self._phase1()
                    with torch.cuda.graph(g2):
                        self._phase2()
                    self.graphs = (g1, g2)
        except Exception as ex:      # capture is an optimisation; eager launches are the same kernels
            self.graphs = None
            self.capture_error = repr(ex)
            torch.cuda.synchronize()
        # undo the warm-up updates
        e.flat.copy_(flat)
        self.acc.zero_()
        if self.sampler is not None:
            self.sampler.counter.zero_()
        for k, v in state.items():
            e.opt_state[k].copy_(v)
        e.pack_weights()
        return self.graphs is not None

    def load(self, x, fut, eps):
        """Stage the next batch.  Device tensors are copied into the static buffers on the current stream.  HOST tensors (pinned)
        go through a copy stream into one of two staging sets, so that the upload of batch i+1 overlaps the compute of batch i
        when the caller issues it before synchronising on step i; run() moves the staged set in with a device-to-device copy."""
        if x.is_cuda:
            self.x.copy_(x, non_blocking=True)
            if self.fut is not None and fut is not None:
                self.fut.copy_(fut, non_blocking=True)
            self.eps.copy_(eps, non_blocking=True)
            return
        if self._stage is None:
            self._stage = [dict(x=torch.empty_like(self.x), fut=None if self.fut is None else torch.empty_like(self.fut),
                                eps=torch.empty_like(self.eps), ready=torch.cuda.Event(), free=None) for _ in range(2)]
            self._copy_stream = torch.cuda.Stream(device=self.eng.device)
            self._staged = []
        k = self._k_load
        st = self._stage[k]
        cs = self._copy_stream
        with torch.cuda.stream(cs):
            if st["free"] is not None:
                cs.wait_event(st["free"])            # run() has moved the previous contents of this set into the static buffers
            st["x"].copy_(x, non_blocking=True)
            if st["fut"] is not None and fut is not None:
                st["fut"].copy_(fut, non_blocking=True)
            st["eps"].copy_(eps, non_blocking=True)
            st["ready"].record(cs)
        self._staged.append((k, fut is not None))
        self._k_load ^= 1

    def _take_staged(self):
        if not self._staged:
            return
        k, has_fut = self._staged.pop(0)
        st = self._stage[k]
        cur = torch.cuda.current_stream()
        cur.wait_event(st["ready"])
        self.x.copy_(st["x"], non_blocking=True)
        if self.fut is not None and has_fut:
            self.fut.copy_(st["fut"], non_blocking=True)
        self.eps.copy_(st["eps"], non_blocking=True)
        if st["free"] is None:
            st["free"] = torch.cuda.Event()
        st["free"].record(cur)

    def run(self):
        """Executes one step on the static buffers; returns the device loss vector."""
        if self.eng._external_dirty:         # load_state_dict / manual weight edits since the last full re-pack
            self.eng.pack_weights()
        if self._staged:
            self._take_staged()
        if self.graphs is not None:
            self.graphs[0].replay()
        else:
            self._phase1()
        if self.world > 1 and not self.nccl_in_graph:
            self._allreduce()
        if self.graphs is not None:
            if self.graphs[1] is not None:
                self.graphs[1].replay()
        else:
            self._phase2()
        # partial re-pack (only the formats this batch size reads): any other consumer re-packs fully (_ensure_packed); an
        # EXTERNAL weight change is tracked separately (Engine._external_dirty) and picked up at the top of run()
        self.eng._packed_version = None
        return self.losses
