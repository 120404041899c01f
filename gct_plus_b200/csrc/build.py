"""Builds libgct_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "libgct_b200.so")
SOURCES = ["gct_abi.cu"]
HEADERS = ["common.cuh", "epilogue.cuh", "elementwise.cuh", "gemm_simt.cuh", "gemm_tc.cuh", "gemm_rownorm.cuh", "attention.cuh", "attention_tc.cuh", "decode.cuh", "decode_attn_mma.cuh", "decode_zattn.cuh",
           "model.cuh", os.path.join("..", "..", "include", "gct_b200.h")]


def _stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(HERE, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = True) -> str:
    if not force and not _stale():
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
           "-shared", "-DGCT_SM_TARGET=100", "-split-compile=0", "-o", OUT] + [os.path.join(HERE, s) for s in SOURCES]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv)
