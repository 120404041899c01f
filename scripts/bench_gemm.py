"""Micro-benchmark of gct_gemm (tcgen05) for given shapes / tile configs.  CUDA events, L2-warm (decode shapes are
L2-resident by nature) and a rotating-buffer mode for the large training shapes."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gct_plus_b200._lib as L  # noqa: E402

dev = torch.device("cuda:0")
lib = L.lib()


def run(M, N, K, bn, stages, split, a_mn=False, b_mn=False, iters=50, nbuf=1, accum=False, flags=0):
    As = [torch.randn((K, M) if a_mn else (M, K), device=dev).bfloat16() for _ in range(nbuf)]
    Bs = [torch.randn((K, N) if b_mn else (N, K), device=dev).bfloat16() for _ in range(nbuf)]
    outT = None if accum else torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    out32 = torch.zeros(M, N, device=dev) if accum else None
    bias = torch.randn(N, device=dev)
    hint = bn + 1000 * stages

    def call(i):
        A, B = As[i % nbuf], Bs[i % nbuf]
        L.check(lib.gct_gemm(L.ptr(A), int(a_mn), A.stride(0), L.ptr(B), int(b_mn), B.stride(0), M, N, K, L.ptr(bias), None, None,
                             None, L.ptr(out32), L.ptr(outT), N, (4 if accum else 0) | flags, split, hint, 1, L.stream_ptr()))
    for i in range(3):
        call(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            call(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    tf = 2.0 * M * N * K / us / 1e6
    print(f"M={M:6d} N={N:5d} K={K:5d} bn={bn:3d} st={stages} split={split} amn={int(a_mn)} bmn={int(b_mn)} nbuf={nbuf}: {us:8.2f} us  {tf:8.1f} TFLOP/s",
          flush=True)


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "decode"
    if mode == "nostore":
        M = 41472
        for (N, K) in [(1536, 512), (512, 512), (2048, 512), (512, 2048)]:
            run(M, N, K, 0, 0, 1, nbuf=3, iters=10)
            run(M, N, K, 0, 0, 1, nbuf=3, iters=10, flags=16)
            run(M, N, K, 0, 0, 1, nbuf=3, iters=10, flags=32)
            run(M, N, K, 0, 0, 1, nbuf=3, iters=10, flags=64)
        sys.exit(0)
    if mode == "train2":
        M = 41472
        for (N, K) in [(1536, 512), (512, 512), (2048, 512), (512, 2048)]:
            for bn in (0, 128, 256, 64):
                run(M, N, K, bn, 0, 1, nbuf=3, iters=10)
        for (N, K) in [(1536, 512), (512, 512), (2048, 512), (512, 2048)]:
            for bn in (0, 128, 256):
                run(4096, N, K, bn, 0, 1, nbuf=1, iters=20)
        for bn in (0, 128, 256):
            run(M, 512, 1536, bn, 0, 1, b_mn=True, nbuf=3, iters=10)
            run(1536, 512, M, bn, 0, 9, a_mn=True, b_mn=True, nbuf=3, iters=10, accum=True)
        sys.exit(0)
    if mode == "decode":
        for (N, K) in [(1536, 512), (512, 512), (2048, 512), (512, 2048)]:
            for bn, stv in [(128, 3), (128, 6), (64, 4), (64, 8), (32, 4), (32, 8), (16, 8)]:
                run(512, N, K, bn, stv, 1)
        for split in (2, 4, 8):
            for bn, stv in [(32, 8), (64, 8), (128, 6)]:
                run(512, 512, 2048, bn, stv, split, accum=True)
                run(512, 512, 512, bn, stv, split, accum=True)
    else:
        M = 41472
        for (N, K) in [(1536, 512), (512, 512), (2048, 512), (512, 2048)]:
            for bn, stv in [(128, 3), (128, 6), (256, 4)]:
                run(M, N, K, bn, stv, 1, nbuf=3, iters=10)
        for bn, stv in [(128, 3), (128, 6), (256, 4)]:
            run(M, 512, 1536, bn, stv, 1, b_mn=True, nbuf=3, iters=10)          # dgrad
            for split in (4, 9, 18):
                run(1536, 512, M, bn, stv, split, a_mn=True, b_mn=True, nbuf=3, iters=10, accum=True)   # wgrad
