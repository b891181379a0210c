"""Blind-debug helper: stage-by-stage comparison of the CUDA path with the CPU oracle (prints, never asserts)."""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import vame_oracle as vo
from vame_b200.engine import Engine

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def rel(a, b):
    a = a.detach().double().cpu() if torch.is_tensor(a) else torch.as_tensor(a).double()
    b = b.detach().double().cpu() if torch.is_tensor(b) else torch.as_tensor(b).double()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def stage(name, fn):
    try:
        fn()
    except Exception:
        print("STAGE %s FAILED" % name)
        traceback.print_exc()
    torch.cuda.synchronize()


def run_case(name, B_override=None):
    g = np.load(os.path.join(GOLD, "step_%s.npz" % name))
    B, T, F, Z, H, fut, S = (int(v) for v in g["cfg"])
    print("==== case", name, (B, T, F, Z, H, fut, S))
    torch.manual_seed(19)
    port = vo.RefPort(2 * T, Z, F, bool(fut), S, hidden=H)
    eng = Engine(F, T, Z, H, H, H, bool(fut), S, False, device="cuda")
    eng.load_state_dict(port.state_dict())
    x, xf, eps = (torch.from_numpy(g[k]) for k in ("x", "fut", "eps"))
    xc, xfc, epsc = x.cuda(), xf.cuda(), eps.cuda()
    hp = dict(beta=1.0, kl_weight=float(g["hp_kl_weight"]), kmeans_loss=Z, kmeans_lambda=0.1, bsize=B)

    def s_enc():
        with torch.no_grad():
            ref = port.encode(x)
        got = eng.encoder_forward(xc)
        print("encoder hidden rel err", rel(got, ref), "pieces", [rel(got[:, i * H:(i + 1) * H], ref[:, i * H:(i + 1) * H]) for i in range(4)])
    stage("encoder", s_enc)

    def s_lam():
        with torch.no_grad():
            hid = port.encode(x)
            zr, mur, lvr = port.lmbda(hid, eps)
        z, mu, lv = eng.lambda_forward(hid.cuda(), epsc)
        print("lambda rel err z/mu/lv", rel(z, zr), rel(mu, mur), rel(lv, lvr))
    stage("lambda", s_lam)

    def s_dec():
        with torch.no_grad():
            zr = torch.from_numpy(g["z"])
            pr = port.decode(zr, "decoder")
        p = eng.decoder_forward(zr.cuda(), 0)
        print("decoder rel err", rel(p, pr))
        if fut:
            with torch.no_grad():
                pf = port.decode(zr, "decoder_future")
            p2 = eng.decoder_forward(zr.cuda(), 1)
            print("decoder_future rel err", rel(p2, pf))
    stage("decoder", s_dec)

    terms, grads, aux = vo.train_step(port, x, xf, eps, hp)

    def s_fwd():
        out = eng.forward(xc, epsc, save=True)
        for k in ("pred", "future", "z", "mu", "logvar"):
            if k in out:
                print("forward", k, "rel err", rel(out[k], aux[k]), "vs golden", rel(out[k], g[k]))
        cfg = eng.loss_cfg(kmeans_loss=Z, kmeans_lambda=0.1, bsize=B, beta=1.0, kl_weight=hp["kl_weight"])
        ls = eng.loss(cfg, xfc if fut else None, want_grads=True).cpu().tolist()
        print("losses got", ls[:5])
        print("losses ref", [terms["rec"], terms.get("fut", 0.0), terms["kl"], terms["kmeans"], terms["total"]])
        eng.backward(cfg)
        torch.cuda.synchronize()
        gv = eng.views(eng.grad)
        worst = 0
        for k in eng.names:
            e = rel(gv[k], grads[k])
            worst = max(worst, e)
            if e > 1e-4:
                print("  grad", k, "rel err", e, "ref max", float(grads[k].abs().max()), "got max", float(gv[k].abs().max()))
        print("worst grad rel err", worst)
    stage("train_step", s_fwd)

    def s_adam():
        opt = vo.make_optimizer(port)
        cfg = eng.loss_cfg(kmeans_loss=Z, kmeans_lambda=0.1, bsize=B, beta=1.0, kl_weight=hp["kl_weight"])
        for it in range(3):
            vo.train_step(port, x, xf, eps, hp, optimizer=opt)
            eng.forward(xc, epsc, save=True)
            eng.loss(cfg, xfc if fut else None, want_grads=True)
            eng.backward(cfg)
            eng.adam_step(lr=5e-4)
        sd = port.state_dict()
        v = eng.views()
        errs = {k: float((v[k].cpu() - sd[k]).abs().max()) for k in eng.names}
        print("adam 3 steps: max abs weight diff", max(errs.values()), "frac>5e-6:",
              float(np.mean([float(((v[k].cpu() - sd[k]).abs() > 5e-6).float().mean()) for k in eng.names])))
    stage("adam", s_adam)


def run_embed():
    g = np.load(os.path.join(GOLD, "embed_synth.npz"))
    T, F, Z, H = (int(v) for v in g["cfg"])
    torch.manual_seed(19)
    port = vo.RefPort(2 * T, Z, F, True, 15, hidden=H)
    eng = Engine(F, T, Z, H, H, H, True, 15, False, device="cuda")
    eng.load_state_dict(port.state_dict())
    series = torch.from_numpy(np.ascontiguousarray(g["series"].T)).float().cuda()
    lat = eng.embed(series, chunk=256)
    print("embed synth shape", tuple(lat.shape), "rel err vs reference", rel(lat, g["latent"]))
    lat2 = eng.embed(series, chunk=128)
    print("embed chunk128 vs chunk256", rel(lat2, lat))


if __name__ == "__main__":
    cases = sys.argv[1:] or ["tiny_fut", "small_nofut", "odd_fut", "c2_h256"]
    for c in cases:
        stage(c, lambda: run_case(c))
    stage("embed", run_embed)
    print("done")
