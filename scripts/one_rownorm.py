"""Residual projection + Norm: the fused kernel (gct_gemm_rownorm) against the GEMM + norm_fwd pair it replaces.
usage: one_rownorm.py M K"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gct_plus_b200._lib as L  # noqa: E402

M, K = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda:0")
lib = L.lib()
A = torch.randn(M, K, device=dev).bfloat16()
W = (torch.randn(512, K, device=dev) / K ** 0.5).bfloat16()
bias, alpha, beta = torch.randn(512, device=dev), torch.ones(512, device=dev), torch.zeros(512, device=dev)
x = torch.randn(M, 512, device=dev)
xn = torch.empty(M, 512, device=dev, dtype=torch.bfloat16)
st = L.stream_ptr()


def fused():
    L.check(lib.gct_gemm_rownorm(L.ptr(A), K, L.ptr(W), K, M, K, L.ptr(bias), L.ptr(x), L.ptr(x), L.ptr(alpha), L.ptr(beta), L.ptr(xn), None, 1e-6, st))


def pair():
    L.check(lib.gct_gemm(L.ptr(A), 0, K, L.ptr(W), 0, K, M, 512, K, L.ptr(bias), L.ptr(x), None, None, L.ptr(x), None, 512, 0, 1, 0, 1, st))
    L.check(lib.gct_norm_fwd(L.ptr(x), L.ptr(alpha), L.ptr(beta), L.ptr(xn), None, M, 512, 1, st))


for name, fn in (("fused, boxed residual", fused), ("gemm + norm_fwd", pair), ("fused, per-row residual", fused), ("fused, boxed residual", fused), ("gemm + norm_fwd", pair)):
    lib.gct_set_rownorm_fusion(1 | (4 if "per-row" in name else 0))
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"M={M} K={K} {name:24s}: {e0.elapsed_time(e1) * 100:.1f} us", flush=True)
