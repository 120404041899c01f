"""CTA-pair (cta_group::2) persistent GEMM against torch fp32 matmul, then timing vs the single-CTA kernel."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gct_plus_b200._lib as L  # noqa: E402
from gpu_common import DEV, gemm  # noqa: E402

lib = L.lib()
torch.manual_seed(0)
for (M, N, K) in [(19584, 512, 512), (19584 - 40, 1536, 512), (41472, 2048, 512), (30000, 512, 2048)]:
    A = torch.randn(M, K, device=DEV).bfloat16()
    B = torch.randn(N, K, device=DEV).bfloat16()
    bias = torch.randn(N, device=DEV)
    res = torch.randn(M, N, device=DEV)
    ref = A.float() @ B.float().t()
    for pair in (2, 0):
        lib.gct_set_cta_pair_gemm(pair)
        _, outT, _ = gemm(A, B, M, N, K, bias=bias, want_T=True)
        e1 = float((outT.float() - (ref + bias)).abs().max() / ref.abs().max())
        out, _, _ = gemm(A, B, M, N, K, bias=bias, res=res)
        e2 = float((out - (ref + bias + res)).abs().max() / ref.abs().max())
        _, g, aux = gemm(A, B, M, N, K, bias=bias, flags=1 | 128, want_T=True)
        e3 = float((g.float() - torch.nn.functional.gelu(ref + bias)).abs().max() / ref.abs().max())
        print(f"M={M} N={N} K={K} pair={pair}: rel err bf16-out {e1:.2e}, fp32+res {e2:.2e}, gelu {e3:.2e}", flush=True)
        assert e1 < 1e-2 and e2 < 2e-3 and e3 < 1e-2
lib.gct_set_cta_pair_gemm(0)
print("pair GEMM ok")
# dgrad (B MN-major) and split-K wgrad (A, B MN-major, accumulate) through the pair kernel (gct_set_cta_pair_gemm(2))
for (M, N, K, a_mn, b_mn, split) in [(41472, 512, 1536, False, True, 1), (1536, 512, 41472, True, True, 9), (512, 512, 40448, True, True, 37)]:
    A = torch.randn(M, K, device=DEV).bfloat16()
    B = torch.randn(N, K, device=DEV).bfloat16()
    ref = A.float() @ B.float().t()
    As = A.t().contiguous() if a_mn else A
    Bs = B.t().contiguous() if b_mn else B
    for pair in (2, 0):
        lib.gct_set_cta_pair_gemm(pair)
        if split == 1:
            out, _, _ = gemm(As, Bs, M, N, K, a_mn=a_mn, b_mn=b_mn)
        else:
            out = torch.ones(M, N, device=DEV)
            gemm(As, Bs, M, N, K, a_mn=a_mn, b_mn=b_mn, flags=4, split_k=split, out32=out)
            out = out - 1.0
        e = float((out - ref).abs().max() / ref.abs().max())
        print(f"M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn} split={split} pair={pair}: rel err {e:.2e}", flush=True)
        assert e < 2e-3
lib.gct_set_cta_pair_gemm(2)
print("pair dgrad / wgrad ok")
