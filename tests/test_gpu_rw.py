"""GPU parity of the resident-weight cluster sweeps (csrc/gru_rw.cu) for every hidden size they are instantiated for
(H = 64, 128, 192, 256 <-> NKC = 1..4), against the CPU oracle and against the slice kernels of csrc/gru.cu on the same
inputs.  Tolerances as in tests/test_gpu_model.py."""
import pytest
import torch

from oracle import vame_oracle as vo

pytestmark = pytest.mark.gpu

from tests.test_gpu_model import GRAD_TOL, OUT_TOL, check_step, make, rel  # noqa: E402


@pytest.fixture()
def lib():
    from vame_b200 import _lib
    L = _lib.lib()
    old = {k: L.vame_get_option(k) for k in (b"rw", b"rw2", b"rw_priv", b"rw_ng")}
    yield L
    for k, v in old.items():
        L.vame_set_option(k, v)


def _step(eng, x, xf, eps, fut, B, Z):
    out = eng.forward(x.cuda(), eps.cuda(), save=True)
    cfg = eng.loss_cfg(kmeans_loss=Z, kmeans_lambda=0.1, bsize=B, beta=1.0, kl_weight=1.0)
    eng.loss(cfg, xf.cuda() if fut else None, want_grads=True)
    eng.backward(cfg)
    torch.cuda.synchronize()
    return {k: v.clone() for k, v in out.items()}, {k: v.clone() for k, v in eng.views(eng.grad).items()}


@pytest.mark.parametrize("H,B,T,F,Z,S,fut,rw2,priv,ng", [
    (64, 48, 12, 10, 8, 5, True, 0, 0, 0),
    (128, 130, 9, 7, 12, 4, True, 0, 0, 0),       # ragged: 130 rows -> B_pad 256, 16 clusters per direction
    (192, 32, 6, 5, 6, 0, False, 0, 0, 0),
    (256, 256, 30, 24, 30, 0, False, 0, 0, 0),    # BASELINE configs[1], cluster-barrier kernels
    (256, 256, 30, 24, 30, 0, False, 1, 0, 0),    # BASELINE configs[1], barrier-free kernels (st.async / mbarrier exchange)
    (256, 256, 30, 24, 30, 0, False, 1, 1, 0),    # ... with the private interchange layouts (the default)
    (256, 96, 30, 24, 30, 15, True, 1, 1, 0),     # decoder + future decoder sweeps side by side
    (256, 96, 30, 24, 30, 15, True, 1, 0, 0),
    (256, 200, 3, 12, 30, 1, True, 1, 1, 0),      # very short sweeps (1 and 3 steps), ragged batch
    # two 16-row groups per cluster (32 rows on the N side of every MMA, single operand buffer, one BPTT A tile in tensor memory)
    (256, 96, 5, 24, 30, 2, True, 1, 1, 2),       # forced at a small batch, decoder + future decoder side by side
    (256, 200, 3, 12, 30, 1, True, 1, 0, 2),      # without the private layouts, ragged batch, 1- and 3-step sweeps
    (256, 512, 30, 24, 30, 0, False, 1, 1, 0),    # BASELINE configs[4] per-GPU batch: chosen automatically (one wave of 32 clusters)
    (256, 512, 12, 60, 50, 6, True, 1, 1, 0),     # BASELINE configs[2] shapes (shorter window)
])
def test_rw_sweeps_vs_oracle_and_slice_kernels(lib, H, B, T, F, Z, S, fut, rw2, priv, ng):
    lib.vame_set_option(b"rw2", rw2)
    lib.vame_set_option(b"rw_priv", priv)
    lib.vame_set_option(b"rw_ng", ng)
    port, eng = make(T, Z, F, fut, S, H)
    x, xf, eps = vo.synthetic_batch(B, T, F, max(S, 1), Z)
    xf = xf[:, :S] if fut else xf
    hp = dict(beta=1.0, kl_weight=1.0, kmeans_loss=Z, kmeans_lambda=0.1, bsize=B)
    t0 = lib.vame_get_option(b"rw_timeouts")
    lib.vame_set_option(b"rw", 3)
    n0 = lib.vame_launch_count()
    check_step(port, eng, x, xf, eps, hp)                      # vs the CPU oracle
    n_rw = lib.vame_launch_count() - n0
    out_rw, g_rw = _step(eng, x, xf, eps, fut, B, Z)
    assert lib.vame_get_option(b"rw_timeouts") == t0, "a bounded wait inside the rw kernels gave up"
    lib.vame_set_option(b"rw", 0)
    n0 = lib.vame_launch_count()
    out_sl, g_sl = _step(eng, x, xf, eps, fut, B, Z)
    n_sl = lib.vame_launch_count() - n0
    assert n_rw < n_sl, "the rw kernels were not used (%d vs %d launches)" % (n_rw, n_sl)
    for k in out_rw:
        assert rel(out_rw[k], out_sl[k]) <= OUT_TOL / 4, k
    for k in g_rw:
        assert rel(g_rw[k], g_sl[k]) <= GRAD_TOL / 2, k


def test_rw_forward_only_and_embed(lib):
    """inference workspaces keep 2 h slots (out_slots = 2): eval forward and the chunked embedding through the rw kernels"""
    T, F, Z, H = 12, 10, 8, 64
    port, eng = make(T, Z, F, True, 5, H)
    x, xf, eps = vo.synthetic_batch(40, T, F, 5, Z)
    outs = {}
    for rw in (3, 0):
        lib.vame_set_option(b"rw", rw)
        o = eng.forward(x.cuda(), None, save=False)
        series = torch.randn(400, F, generator=torch.Generator().manual_seed(5)).cuda()
        outs[rw] = (o["pred"].clone(), o["mu"].clone(), eng.embed(series, chunk=128).clone())
    for a, b in zip(outs[3], outs[0]):
        assert rel(a, b) <= OUT_TOL / 4
    with torch.no_grad():
        hid = port.encode(x)
        _, mur, _ = port.lmbda(hid, None)
    assert rel(outs[3][1], mur) <= OUT_TOL
