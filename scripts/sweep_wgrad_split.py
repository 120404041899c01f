"""Split-K sweep of the weight-gradient GEMMs (dW[Nout, Kin] += dY^T X over R = 41 472 or 51 712 token rows; both operands MN-major,
fp32 atomics into dW) at the model's shapes.  The model's rule (wgrad_split in csrc/model.cuh) is marked with *."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gct_plus_b200._lib as L  # noqa: E402
from bench_gemm import run  # noqa: E402


def rule_r1(Mout, Nout, R):          # the rule until late round 2: two waves of single CTAs
    tiles = -(-Mout // 128) * -(-Nout // 256)
    s = (2 * 148 + tiles - 1) // tiles
    return max(1, min(s, (-(-R // 64)) // 2))


def rule(Mout, Nout, R):             # wgrad_split in csrc/model.cuh
    tiles = -(-Mout // 256) * -(-Nout // 256)
    s = (72 if tiles <= 8 else 144) // tiles
    return max(1, min(s, (-(-R // 64)) // 2))


R = int(sys.argv[1]) if len(sys.argv) > 1 else 41472
for (Mout, Nout) in [(512, 512), (1024, 512), (1536, 512), (2048, 512), (512, 2048), (256, 512), (512, 128)]:
    r = rule(Mout, Nout, R)
    for split in sorted({r, rule_r1(Mout, Nout, R), max(1, r // 2), r + r // 2}):
        print("*" if split == r else " ", end="")
        run(Mout, Nout, R, 0, 0, split, a_mn=True, b_mn=True, iters=20, nbuf=2, accum=True)
