"""Helpers shared by the -m gpu tests: model construction from golden fixtures, raw op calls."""
import ctypes as C

import torch

from helpers import O, cfg_from_fixture, load_golden, sd_from_fixture  # noqa: F401
import gct_plus_b200._lib as L
from gct_plus_b200.Model import Cvaetf, Vaetf

DEV = "cuda:0"


def build_model(fx, dtype, dropout=0.1, sd=None):
    a = fx["arch"]
    cls = Vaetf if fx["model_type"] == "vaetf" else Cvaetf
    m = cls(32, 32, N=a["N"], d_model=a["d_model"], dff=a["dff"], h=a["h"], latent_dim=a["latent_dim"], dropout=dropout,
            nconds=fx["nconds"], use_cond2dec=fx.get("use_cond2dec", False), use_cond2lat=fx.get("use_cond2lat", False),
            compute_dtype=dtype)
    sd = sd if sd is not None else sd_from_fixture(fx)
    m.load_state_dict(sd)
    return m.to(DEV), sd


def gemm(A, B, M, N, K, *, a_mn=False, b_mn=False, bias=None, res=None, flags=0, split_k=1, bn=0, dtype="bf16",
         out32=None, want_T=False, aux_in=None):
    """C = A(m,k) B(n,k) through gct_gemm; A/B are 2-D tensors in their storage layout."""
    lib = L.lib()
    dt = L.DTYPE_BF16 if dtype == "bf16" else L.DTYPE_F32
    tdt = torch.bfloat16 if dtype == "bf16" else torch.float32
    if out32 is None and not want_T:
        out32 = torch.zeros((M, N), device=A.device, dtype=torch.float32)
    outT = torch.zeros((M, N), device=A.device, dtype=tdt) if want_T else None
    aux_out = torch.zeros((M, N), device=A.device, dtype=tdt) if (flags & 1) else None
    L.check(lib.gct_gemm(L.ptr(A), int(a_mn), A.stride(0), L.ptr(B), int(b_mn), B.stride(0), M, N, K, L.ptr(bias), L.ptr(res),
                         L.ptr(aux_in), L.ptr(aux_out), L.ptr(out32), L.ptr(outT), N, flags, split_k, bn, dt, L.stream_ptr()),
            "gct_gemm")
    torch.cuda.synchronize()
    return out32, outT, aux_out
