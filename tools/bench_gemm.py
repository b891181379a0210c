"""Event-timed micro-benchmark of the tcgen05 P16 GEMM on the shapes the train step uses (C2: rows = T*B_pad = 7680).
Inputs are re-used across iterations (L2-warm, as in the step where the operands were just produced)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_gpu_gemm import pack
from vame_b200 import _lib as L

lib = L.lib()
torch.manual_seed(0)
SHAPES = [
    # (M, N, K, accumulate, splits, label)
    (768, 256, 7680, 1, 12, "dW_hh / dW_ih(l1) split-K 12"),
    (768, 256, 7680, 1, 9, "dW_hh split-K 9 (116 SM cap)"),
    (768, 256, 7680, 1, 6, "dW_hh split-K 6"),
    (768, 256, 7680, 0, 1, "same shape, no split, plain store"),
    (768, 24, 7680, 1, 24, "dW_ih(l0) split-K 24"),
    (7680, 1536, 512, 0, 1, "gi1 (feature-major in the step; row-major here)"),
    (7680, 512, 1536, 0, 1, "dx1"),
    (7680, 1536, 64, 0, 1, "gi0"),
]
for (M, N, K, acc, splits, label) in SHAPES:
    a = torch.randn(M, K, device="cuda")
    b = torch.randn(N, K, device="cuda")
    ap, bp = pack(a), pack(b)
    c = torch.zeros(M, N, device="cuda")
    nkc = (K + 63) // 64
    st = L.cur_stream()
    for _ in range(3):
        lib.vame_gemm_p16(L.ptr(ap), nkc, L.ptr(bp), nkc, M, N, L.ptr(c), N, None, acc, splits, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    for _ in range(n):
        lib.vame_gemm_p16(L.ptr(ap), nkc, L.ptr(bp), nkc, M, N, L.ptr(c), N, None, acc, splits, st)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    flops = 2.0 * M * N * K
    c.zero_()
    lib.vame_gemm_p16(L.ptr(ap), nkc, L.ptr(bp), nkc, M, N, L.ptr(c), N, None, acc, splits, st)
    torch.cuda.synchronize()
    err = (c.double() - a.double() @ b.double().T).abs().max().item()
    print("%-52s M=%5d N=%5d K=%5d  %7.2f us  %6.1f TFLOP/s (fp32-equivalent)  max err %.2e" % (label, M, N, K, us, flops / us * 1e-6, err), flush=True)
