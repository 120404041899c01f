"""A/B of library switches on the cfg-3 training step inside one process (CUDA events, 10 steps each, interleaved twice)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import gct_plus_b200._lib as L  # noqa: E402
from gct_plus_b200.Model import Cvaetf  # noqa: E402
from gct_plus_b200.Train.trainer1 import FusedTrainer  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
lib = L.lib()
model = Cvaetf(32, 32, dropout=0.1, nconds=3, use_cond2lat=True, compute_dtype="bf16", **bench.ARCH).to(dev).train()
tr = FusedTrainer(model, "pvaetf")
batch = bench.make_train_batch(512, 78, 3, 0, 1, dev=dev)
settings = [("pair=0 ew4=1", 0, 1), ("pair=1 ew4=1", 1, 1), ("pair=2 ew4=1", 2, 1), ("pair=2 ew4=0", 2, 0)]


def run(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        tr.step(batch, 0.5)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for _ in range(3):
    tr.step(batch, 0.5)
for rep in range(2):
    for name, pair, ew4 in settings:
        lib.gct_set_cta_pair_gemm(pair)
        lib.gct_set_epilogue_warps16(ew4)
        run(2)
        print(f"rep {rep} {name}: {run(10):.3f} ms/step", flush=True)
