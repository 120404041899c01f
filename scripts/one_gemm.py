"""Runs one tcgen05 GEMM shape a few times (target for `ncu --set full`)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gct_plus_b200._lib as L  # noqa: E402

M, N, K, bn = (int(x) for x in sys.argv[1:5])
dev = torch.device("cuda:0")
lib = L.lib()
A = torch.randn(M, K, device=dev).bfloat16()
B = torch.randn(N, K, device=dev).bfloat16()
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
bias = torch.randn(N, device=dev)
for _ in range(4):
    L.check(lib.gct_gemm(L.ptr(A), 0, K, L.ptr(B), 0, K, M, N, K, L.ptr(bias), None, None, None, None, L.ptr(out), N, 0, 1, bn, 1,
                         L.stream_ptr()))
torch.cuda.synchronize()
print("ok")
