"""GPU parity of the packing kernel and the tcgen05 3-pass bf16-split GEMM vs a plain PyTorch fp32/fp64 reference."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
    from vame_b200 import _lib as L
    return L


def unpack_p16(buf, R, K, RB):
    """Python inverse of the P16 layout (hi + lo planes) -> float32 [R, K]."""
    nrb, nkc = (R + RB - 1) // RB, (K + 63) // 64
    t = buf.view(torch.bfloat16).reshape(nrb, nkc, 2, RB // 8, 8, 8, 8).float()
    x = t[:, :, 0] + t[:, :, 1]                       # [nrb, nkc, r8, k8, 8, 8]
    x = x.permute(0, 2, 4, 1, 3, 5).reshape(nrb * RB, nkc * 64)
    return x[:R, :K]


def pack(x, RB=128, transposed=False, R=None, K=None):
    L = _lib()
    lib = L.lib()
    if R is None:
        R, K = (x.shape[1], x.shape[0]) if transposed else x.shape
    nbytes = lib.vame_p16_bytes(R, K, RB)
    out = torch.zeros(nbytes, dtype=torch.uint8, device=x.device)
    rs, ks = (x.shape[1], x.shape[0]) if transposed else x.shape
    L.check(lib.vame_pack_p16(L.ptr(x), x.stride(0), int(transposed), R, K, rs, ks, None, None, RB, L.ptr(out), L.cur_stream()),
            "pack")
    return out


@pytest.mark.parametrize("R,K,RB", [(128, 64, 128), (200, 100, 128), (96, 256, 96), (257, 130, 128)])
def test_pack_roundtrip(R, K, RB):
    torch.manual_seed(0)
    x = torch.randn(R, K, device="cuda")
    p = pack(x, RB)
    torch.cuda.synchronize()
    y = unpack_p16(p, R, K, RB)
    assert (y - x).abs().max().item() <= 2.0 ** -15 * x.abs().max().item()
    xt = torch.randn(K, R, device="cuda")
    pt = pack(xt, RB, transposed=True)
    yt = unpack_p16(pt, R, K, RB)
    assert (yt - xt.T).abs().max().item() <= 2.0 ** -15 * xt.abs().max().item()


@pytest.mark.parametrize("M,N,K,splits", [(128, 128, 64, 1), (128, 128, 256, 1), (256, 384, 512, 1), (300, 200, 100, 1),
                                           (768, 256, 4096, 8), (1000, 72, 24, 1)])
def test_gemm_matches_fp32(M, N, K, splits):
    L = _lib()
    lib = L.lib()
    torch.manual_seed(1)
    a = torch.randn(M, K, device="cuda")
    b = torch.randn(N, K, device="cuda") * 0.1
    bias = torch.randn(N, device="cuda")
    ap, bp = pack(a), pack(b)
    nkc = (K + 63) // 64
    c = torch.zeros(M, N, device="cuda")
    L.check(lib.vame_gemm_p16(L.ptr(ap), nkc, L.ptr(bp), nkc, M, N, L.ptr(c), N, L.ptr(bias), int(splits > 1), splits,
                              L.cur_stream()), "gemm")
    torch.cuda.synchronize()
    ref64 = a.double() @ b.double().T + bias.double()
    err = (c.double() - ref64).abs().max().item()
    scale = ref64.abs().max().item()
    ref32 = torch.addmm(bias, a, b.T)
    err32 = (ref32.double() - ref64).abs().max().item()
    # fp32-accurate: within 2e-5 of the fp64 value relative to the output scale (the torch fp32 GEMM itself is ~1e-6)
    assert err <= 2e-5 * scale, (err, err32, scale)
