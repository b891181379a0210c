"""Blind-debug helper: run the tcgen05 GEMM on simple patterns and dump error structure."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_gpu_gemm import pack, unpack_p16
from vame_b200 import _lib as L

lib = L.lib()
torch.manual_seed(0)
for (M, N, K) in [(128, 128, 64), (128, 128, 128), (256, 256, 256)]:
    a = torch.randn(M, K, device="cuda")
    b = torch.randn(N, K, device="cuda")
    ap, bp = pack(a), pack(b)
    au = unpack_p16(ap, M, K, 128)
    print("pack err", (au - a).abs().max().item())
    c = torch.full((M, N), -7.0, device="cuda")
    nkc = (K + 63) // 64
    rc = lib.vame_gemm_p16(L.ptr(ap), nkc, L.ptr(bp), nkc, M, N, L.ptr(c), N, None, 0, 1, L.cur_stream())
    torch.cuda.synchronize()
    ref = a.double() @ b.double().T
    err = (c.double() - ref).abs()
    print("MNK", M, N, K, "rc", rc, "max err", err.max().item(), "ref scale", ref.abs().max().item())
    if err.max().item() > 1e-3:
        bad = (err > 1e-3)
        print("  bad frac", bad.float().mean().item(), "bad rows", bad.any(1).sum().item(), "bad cols", bad.any(0).sum().item())
        print("  c[0,:8]", c[0, :8].tolist())
        print("  r[0,:8]", ref[0, :8].tolist())
        print("  c[:8,0]", c[:8, 0].tolist())
        print("  r[:8,0]", ref[:8, 0].tolist())
        # try to identify permutation: is c == ref with K restricted?
        for kk in (16, 32, 64):
            r2 = a[:, :kk].double() @ b[:, :kk].double().T
            print("  partial K", kk, (c.double() - r2).abs().max().item())
print("done")
