"""Host-side mirror of vame/model/rnn_model.py: same class names, constructor signatures, attribute names,
parameter registration order and state_dict keys as the reference (Encoder :23-45, Lambda :48-76, Decoder :79-109,
Decoder_Future :112-144, RNN_VAE :147-179), but every forward/backward runs in the sm_100a CUDA library.

The torch ``nn.GRU`` / ``nn.Linear`` objects below are PARAMETER CONTAINERS only (they give the reference's default
initialisation and key names, e.g. ``encoder.encoder_rnn.weight_hh_l0_reverse``); they are never called.  After
``.cuda()`` all parameters become views into one flat fp32 buffer owned by ``vame_b200.engine.Engine``.
There is no CPU execution path: calling a module that has not been moved to a CUDA device raises.
"""
import torch
from torch import nn

from ._lib import VameB200Error
from .engine import Engine


import os as _os
_CHECK_DECODER_INPUT = _os.environ.get("VAME_B200_CHECK_DECODER_INPUT", "1") != "0"


def _bigru(in_size, hidden, layers, dropout):
    """Parameter container with the reference's nn.GRU configuration (bias, batch_first, bidirectional)."""
    return nn.GRU(in_size, hidden, layers, bias=True, batch_first=True, dropout=dropout, bidirectional=True)


def _no_grad_needed(*tensors):
    return not (torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors))


class Encoder(nn.Module):
    def __init__(self, NUM_FEATURES, hidden_size_layer_1, hidden_size_layer_2, dropout_encoder):
        super(Encoder, self).__init__()
        self.input_size, self.hidden_size, self.hidden_size_2 = NUM_FEATURES, hidden_size_layer_1, hidden_size_layer_2
        self.n_layers, self.dropout, self.bidirectional, self.hidden_factor = 2, dropout_encoder, True, 4
        self.encoder_rnn = _bigru(NUM_FEATURES, hidden_size_layer_1, 2, dropout_encoder)     # rnn_model.py:34-35
        self._owner = None

    def forward(self, inputs):
        """(B, T, F) -> cat(h_n[0..3]) (B, 4H)   [rnn_model.py:40-45]"""
        eng = _engine_of(self)
        if not _no_grad_needed(inputs):
            raise NotImplementedError("vame_b200: autograd through a stand-alone Encoder is not supported; call the RNN_VAE")
        return eng.encoder_forward(inputs.detach().float())


class Lambda(nn.Module):
    def __init__(self, ZDIMS, hidden_size_layer_1, hidden_size_layer_2, softplus):
        super(Lambda, self).__init__()
        self.hid_dim, self.latent_length, self.softplus = hidden_size_layer_1 * 4, ZDIMS, softplus
        self.hidden_to_mean = nn.Linear(self.hid_dim, ZDIMS)          # rnn_model.py:56
        self.hidden_to_logvar = nn.Linear(self.hid_dim, ZDIMS)        # rnn_model.py:57
        if self.softplus == True:  # noqa: E712  (mirrors the reference's comparison)
            print("Using a softplus activation to ensures that the variance is parameterized as non-negative and activated by a smooth function")
        self._owner = None

    def forward(self, hidden):
        """-> (z, mean, logvar); z = eps*exp(0.5 logvar) + mean in training mode, mean in eval mode [rnn_model.py:63-76]"""
        eng = _engine_of(self)
        if not _no_grad_needed(hidden):
            raise NotImplementedError("vame_b200: autograd through a stand-alone Lambda is not supported; call the RNN_VAE")
        eps = None
        if self.training:
            eps = torch.randn(hidden.shape[0], self.latent_length, device=hidden.device, dtype=torch.float32)
        z, mu, lv = eng.lambda_forward(hidden.detach().float(), eps)
        self.mean, self.logvar = mu, lv
        return z, mu, lv


class Decoder(nn.Module):
    def __init__(self, TEMPORAL_WINDOW, ZDIMS, NUM_FEATURES, hidden_size_rec, dropout_rec):
        super(Decoder, self).__init__()
        self._common(TEMPORAL_WINDOW, ZDIMS, NUM_FEATURES, hidden_size_rec, dropout_rec)
        self.rnn_rec = _bigru(ZDIMS, hidden_size_rec, 1, dropout_rec)                       # rnn_model.py:91-92
        self._heads(ZDIMS, NUM_FEATURES, hidden_size_rec)
        self._which = 0

    def _common(self, T, Z, F, H, dropout):
        self.num_features, self.sequence_length, self.hidden_size, self.latent_length = F, T, H, Z
        self.n_layers, self.dropout, self.bidirectional, self.hidden_factor = 1, dropout, True, 2
        self._owner = None

    def _heads(self, Z, F, H):
        self.latent_to_hidden = nn.Linear(Z, 2 * H)                   # rnn_model.py:96 / :130
        self.hidden_to_output = nn.Linear(2 * H, F)                   # rnn_model.py:97 / :131

    def forward(self, inputs, z):
        """inputs must be z repeated over time, as every reference call site builds it
        (rnn_model.py:169-170, generative_functions.py:35-36,59-60,78-79,97-98) [rnn_model.py:99-109]"""
        eng = _engine_of(self)
        if not _no_grad_needed(inputs, z):
            raise NotImplementedError("vame_b200: autograd through a stand-alone Decoder is not supported; call the RNN_VAE")
        steps = self.sequence_length if self._which == 0 else self.future_steps
        if inputs.shape[0] != z.shape[0] or inputs.shape[1] < steps or inputs.shape[2] != z.shape[1]:
            raise ValueError("decoder inputs must be z repeated over >= %d time steps, got %s" % (steps, tuple(inputs.shape)))
        if _CHECK_DECODER_INPUT and not torch.equal(inputs[:, :steps].float(), z.float().unsqueeze(1).expand(-1, steps, -1)):
            # the library projects z ONCE per sample (the reference's input is z at every step, rnn_model.py:169-170); any other
            # `inputs` would silently be ignored - VAME_B200_CHECK_DECODER_INPUT=1 (default) turns that into an error
            raise ValueError("vame_b200 Decoder.forward: `inputs` is not z repeated over time; only that form is supported")
        return eng.decoder_forward(z.detach().float(), self._which)


class Decoder_Future(Decoder):
    def __init__(self, TEMPORAL_WINDOW, ZDIMS, NUM_FEATURES, FUTURE_STEPS, hidden_size_pred, dropout_pred):
        nn.Module.__init__(self)
        self._common(TEMPORAL_WINDOW, ZDIMS, NUM_FEATURES, hidden_size_pred, dropout_pred)
        self.future_steps = FUTURE_STEPS
        self.rnn_pred = _bigru(ZDIMS, hidden_size_pred, 1, dropout_pred)                    # rnn_model.py:125-126
        self._heads(ZDIMS, NUM_FEATURES, hidden_size_pred)
        self._which = 1


def _engine_of(sub):
    owner = sub._owner() if callable(sub._owner) else None
    if owner is None or owner._engine is None:
        raise VameB200Error("vame_b200: the model is not on a CUDA device (call .cuda()); there is no CPU execution path")
    owner._sync_engine()
    return owner._engine


class _RNNVAEFunction(torch.autograd.Function):
    """Whole-model autograd node: forward and backward both run in the CUDA library.

    The saved activations live in the engine's training workspace of this batch size (one set per batch size, 1.2 GB at
    B = 256), not in ``ctx``: the node records the workspace generation and ``backward`` refuses to run when a later saving
    forward of the same batch size has overwritten it (two forwards before one backward) instead of returning wrong gradients.
    Evaluation / no-grad forwards use a separate workspace and never disturb it."""

    @staticmethod
    def forward(ctx, owner, x, eps, *params):
        eng = owner._engine
        out = eng.forward(x, eps, save=True)
        ctx.owner = owner
        ctx.batch = int(x.shape[0])
        ctx.gen = eng._train_gen[ctx.batch]
        ctx.has_future = "future" in out
        res = (out["pred"],) + ((out["future"],) if ctx.has_future else ()) + (out["z"], out["mu"], out["logvar"])
        return res

    @staticmethod
    def backward(ctx, *grads):
        eng = ctx.owner._engine
        if eng is None or eng._train_gen.get(ctx.batch) != ctx.gen:
            raise RuntimeError("vame_b200: the activations saved by this forward were overwritten by a later forward of the same batch "
                               "size (the engine keeps ONE set of saved activations per batch size); call backward() before the next "
                               "training forward, or run the second forward under torch.no_grad()")
        if ctx.has_future:
            dpred, dfut, dz, dmu, dlv = grads
        else:
            dpred, dz, dmu, dlv = grads
            dfut = None
        g = eng.backward(cfg=None, use_loss_grads=False, dpred=dpred, dfuture=dfut, dz=dz, dmu=dmu, dlogvar=dlv, batch=ctx.batch)
        views = eng.views(g.clone())
        need = ctx.needs_input_grad[3:]
        return (None, None, None) + tuple(views[n] if need[i] else None for i, n in enumerate(eng.names))


class RNN_VAE(nn.Module):
    def __init__(self, TEMPORAL_WINDOW, ZDIMS, NUM_FEATURES, FUTURE_DECODER, FUTURE_STEPS, hidden_size_layer_1,
                 hidden_size_layer_2, hidden_size_rec, hidden_size_pred, dropout_encoder,
                 dropout_rec, dropout_pred, softplus):
        super(RNN_VAE, self).__init__()
        self.FUTURE_DECODER = FUTURE_DECODER
        self.seq_len = int(TEMPORAL_WINDOW / 2)
        self.encoder = Encoder(NUM_FEATURES, hidden_size_layer_1, hidden_size_layer_2, dropout_encoder)
        self.lmbda = Lambda(ZDIMS, hidden_size_layer_1, hidden_size_layer_2, softplus)
        self.decoder = Decoder(self.seq_len, ZDIMS, NUM_FEATURES, hidden_size_rec, dropout_rec)
        if FUTURE_DECODER:
            self.decoder_future = Decoder_Future(self.seq_len, ZDIMS, NUM_FEATURES, FUTURE_STEPS, hidden_size_pred, dropout_pred)
        if dropout_encoder:
            raise NotImplementedError("vame_b200: dropout_encoder > 0 is not implemented (reference default is 0)")
        self._cfg = dict(num_features=NUM_FEATURES, time_window=self.seq_len, zdims=ZDIMS, hidden_enc=hidden_size_layer_1,
                         hidden_rec=hidden_size_rec, hidden_pred=hidden_size_pred, future_decoder=bool(FUTURE_DECODER),
                         future_steps=FUTURE_STEPS if FUTURE_DECODER else 0, softplus=bool(softplus))
        self._engine = None
        import weakref
        ref = weakref.ref(self)
        for m in (self.encoder, self.lmbda, self.decoder) + ((self.decoder_future,) if FUTURE_DECODER else ()):
            object.__setattr__(m, "_owner", ref)

    # ---- device placement: parameters become views of the engine's flat buffer -------------------------
    def _apply(self, fn, recurse=True):
        probe = fn(torch.empty(0, dtype=torch.float32))
        if probe.dtype != torch.float32:
            raise NotImplementedError("vame_b200: only float32 parameters are supported")
        if probe.device.type == "cuda":
            self._bind(probe.device)
            return self
        if self._engine is not None:          # moving back to the host: materialise ordinary tensors
            sd = {k: v.detach().cpu().clone() for k, v in self._engine.views().items()}
            self._engine = None
            for (name, p) in self.named_parameters():
                p.data = sd[name]
                p.grad = None
            return self
        return super()._apply(fn, recurse)

    def _bind(self, device):
        if self._engine is not None and self._engine.device == torch.device(device):
            return
        current = {k: p.detach().clone() for k, p in self.named_parameters()}
        eng = Engine(device=device, **self._cfg)
        assert list(current.keys()) == eng.names, "parameter order differs from the reference state_dict order"
        eng.load_state_dict(current)
        views, gviews = eng.views(), eng.views(eng.grad)
        for name, p in self.named_parameters():
            p.data = views[name]
            p.grad = None
        # p.data = view gives every parameter its OWN version counter: optimizer.step() / p.copy_() never touch flat._version,
        # so the engine watches the parameters' counters to know when its packed tensor-core copies are stale
        eng.watch([p for _, p in self.named_parameters()])
        self._grad_views = gviews
        self._engine = eng
        self.__dict__.pop("_b200_steps", None)      # captured train-step graphs belong to the previous engine

    def _sync_engine(self):
        """Parameters alias engine.flat; the engine compares their version counters (Engine.watch) before every forward and
        re-packs the tensor-core copies when any of them changed (optimizer.step, load_state_dict, p.copy_).  Edits through
        ``p.data`` bump no counter at all: with autograd enabled the forward therefore always re-packs (30 us)."""
        return self._engine

    def load_state_dict(self, state_dict, strict=True, **kw):
        r = super().load_state_dict(state_dict, strict=strict, **kw)
        if self._engine is not None:
            self._engine.mark_dirty()
        return r

    @property
    def engine(self):
        if self._engine is None:
            raise VameB200Error("vame_b200: the model is not on a CUDA device (call .cuda()); there is no CPU execution path")
        return self._engine

    def bind_flat_grads(self):
        """Point every parameter's .grad at its slice of the engine's flat gradient buffer (used by the fused train step)."""
        for name, p in self.named_parameters():
            p.grad = self._grad_views[name]

    def forward(self, seq):
        """-> (prediction, [future,] z, mu, logvar)   [rnn_model.py:162-179]"""
        eng = self.engine
        x = seq.float()
        B = x.shape[0]
        eps = torch.randn(B, self._cfg["zdims"], device=x.device, dtype=torch.float32) if self.training else None
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if need_grad:
            eng._ensure_packed(force=True)           # (see _sync_engine: p.data edits are invisible to version counters)
            out = _RNNVAEFunction.apply(self, x.detach(), eps, *list(self.parameters()))
            if self.FUTURE_DECODER:
                pred, fut, z, mu, lv = out
            else:
                pred, z, mu, lv = out
                fut = None
        else:
            o = eng.forward(x.detach(), eps, save=False)
            pred, fut, z, mu, lv = o["pred"], o.get("future"), o["z"], o["mu"], o["logvar"]
        self.lmbda.mean, self.lmbda.logvar = mu, lv            # side effect of Lambda.forward (rnn_model.py:65-69)
        if self.FUTURE_DECODER:
            return pred, fut, z, mu, lv
        return pred, z, mu, lv
