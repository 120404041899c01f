"""Runs one tcgen05 GEMM shape a few times (target for `ncu --set full`).
usage: one_gemm.py M N K bn [flags [b_mn [res]]]   flags: 1 GELU (+128 saves keep*gelu'), 256 multiply by aux_in;
res = 1: fp32 output with bias + fp32 residual (the out-projection / FFN2 epilogue)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gct_plus_b200._lib as L  # noqa: E402

M, N, K, bn = (int(x) for x in sys.argv[1:5])
flags = int(sys.argv[5]) if len(sys.argv) > 5 else 0
b_mn = int(sys.argv[6]) if len(sys.argv) > 6 else 0
res = int(sys.argv[7]) if len(sys.argv) > 7 else 0
dev = torch.device("cuda:0")
lib = L.lib()
if os.environ.get("GCT_PAIR") is not None:
    lib.gct_set_cta_pair_gemm(int(os.environ["GCT_PAIR"]))
if os.environ.get("GCT_RES_BOX") is not None:
    lib.gct_set_residual_box(int(os.environ["GCT_RES_BOX"]))
if os.environ.get("GCT_EW4") is not None:
    lib.gct_set_epilogue_warps16(int(os.environ["GCT_EW4"]))
A = torch.randn(M, K, device=dev).bfloat16()
B = (torch.randn(K, N, device=dev) if b_mn else torch.randn(N, K, device=dev)).bfloat16()
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if not res else None
out32 = torch.empty(M, N, device=dev) if res else None
res32 = torch.randn(M, N, device=dev) if res == 1 else None
aux_out = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if flags & 1 else None
aux_in = torch.randn(M, N, device=dev).bfloat16() if flags & (2 | 256) else None
bias = torch.randn(N, device=dev) if not (flags & (2 | 256)) else None
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(6):
    if i == 2:
        e0.record()
    L.check(lib.gct_gemm(L.ptr(A), 0, K, L.ptr(B), b_mn, B.stride(0), M, N, K, L.ptr(bias), L.ptr(res32), L.ptr(aux_in), L.ptr(aux_out), L.ptr(out32),
                         L.ptr(out), N, flags, 1, bn, 1, L.stream_ptr()))
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / 4
print(f"ok M={M} N={N} K={K} flags={flags} b_mn={b_mn} res={res}: {us:.1f} us/launch, {2.0 * M * N * K / us / 1e6:.0f} TFLOP/s")
