"""Tile choice of the N = 512 projections with an fp32 residual (out-projections, FFN2, the QKV dgrad): the dispatcher's pick
(bn hint 0) against forced 128-column tiles (single-CTA persistent kernel) and 256-column tiles (CTA-pair kernel)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gct_plus_b200._lib as L  # noqa: E402

dev = torch.device("cuda:0")
lib = L.lib()


def make(M, N, K, bn, b_mn):
    A = torch.randn(M, K, device=dev).bfloat16()
    B = (torch.randn(K, N, device=dev) if b_mn else torch.randn(N, K, device=dev)).bfloat16()
    out32 = torch.empty(M, N, device=dev)
    res32 = torch.randn(M, N, device=dev)
    bias = torch.randn(N, device=dev)
    keep = (A, B, out32, res32, bias)

    def run():
        L.check(lib.gct_gemm(L.ptr(A), 0, K, L.ptr(B), b_mn, B.stride(0), M, N, K, L.ptr(bias), L.ptr(res32), None, None, L.ptr(out32),
                             None, N, 0, 1, bn, 1, L.stream_ptr()))
    return run, keep


def timeit(run, n=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        run()
    e0.record()
    for _ in range(n):
        run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


for M in (30000, 41472, 51712):
    for K, b_mn in ((512, 0), (1024, 0), (2048, 0), (1536, 1)):
        for bn in (0, 128, 256):
            run, keep = make(M, 512, K, bn, b_mn)
            us = min(timeit(run), timeit(run))
            print(f"M={M} N=512 K={K} b_mn={b_mn} bn hint {bn:3d}: {us:7.1f} us  {2.0 * M * 512 * K / us / 1e6:6.0f} TFLOP/s", flush=True)
            del run, keep
