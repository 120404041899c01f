"""A/B inside one process: persistent-GEMM epilogue with the TMA store's wait at the store (gct_set_tma_store(2)) vs deferred to the
staging tile's next write (1, default).  Shapes: the decode (M = 30000) and training (M = 41472) GEMMs of the model."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gct_plus_b200._lib as L  # noqa: E402

dev = torch.device("cuda:0")
lib = L.lib()


def make(M, N, K, flags, res):
    A = torch.randn(M, K, device=dev).bfloat16()
    B = torch.randn(N, K, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if not res else None
    out32 = torch.empty(M, N, device=dev) if res else None
    res32 = torch.randn(M, N, device=dev) if res == 1 else None
    aux_out = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if flags & 128 else None
    bias = torch.randn(N, device=dev)
    keep = (A, B, out, out32, res32, aux_out, bias)

    def run():
        L.check(lib.gct_gemm(L.ptr(A), 0, K, L.ptr(B), 0, K, M, N, K, L.ptr(bias), L.ptr(res32), None, L.ptr(aux_out), L.ptr(out32),
                             L.ptr(out), N, flags, 1, 256, 1, L.stream_ptr()))
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


cases = [("decode QKV", 30000, 1536, 512, 0, 0), ("decode zq", 30000, 1024, 512, 0, 0), ("decode FFN1 GELU", 30000, 2048, 512, 1, 0),
         ("decode FFN2 fp32 res", 30000, 512, 2048, 0, 1), ("train QKV", 41472, 1536, 512, 0, 0), ("train FFN1 GELU+grad", 41472, 2048, 512, 129, 0),
         ("train FFN1 plain", 41472, 2048, 512, 0, 0)]
for name, M, N, K, flags, res in cases:
    run, keep = make(M, N, K, flags, res)
    for rep in range(2):
        for mode, label in ((2, "wait at the store"), (1, "deferred wait")):
            lib.gct_set_tma_store(mode)
            us = timeit(run)
            print(f"{name:22s} M={M} N={N} K={K} rep {rep} {label:18s}: {us:7.1f} us  {2.0 * M * N * K / us / 1e6:6.0f} TFLOP/s", flush=True)
    del run, keep
lib.gct_set_tma_store(1)
