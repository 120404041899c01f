"""usage: ab_train.py [cfg3|cfg4].  A/B of library switches on the training step inside one process (CUDA events, 10 steps each, interleaved twice)."""
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
case = bench.TRAIN_CASES[sys.argv[1] if len(sys.argv) > 1 else "cfg3"]
model = Cvaetf(32, 32, dropout=0.1, nconds=3, use_cond2lat=True, compute_dtype="bf16", **bench.ARCH).to(dev).train()
tr = FusedTrainer(model, case["mt"])
batch = bench.make_train_batch(512, case["S"], 3, case["sca"], 1, dev=dev)
# gct_set_residual_box: bit 0 = GEMM epilogue operands (fp32 residual, multiply-by-aux factor) as TMA boxes; bit 1 set = attention
# kernels store / load per lane instead of as TMA boxes
# third field: gct_set_attention_persistent (bit 0 forward, bit 1 backward; the persistent kernels need the attention boxes)
settings = [("gemm boxes off, attention boxes off", 2, 0), ("gemm boxes on, attention boxes off", 3, 0),
            ("gemm boxes on, attention boxes on, one tile per CTA", 1, 0), ("gemm boxes on, attention boxes on, persistent attention backward", 1, 2),
            ("gemm boxes on, attention boxes on, persistent attention forward + backward", 1, 3)]


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
    for name, mode, persist in settings:
        lib.gct_set_residual_box(mode)
        lib.gct_set_attention_persistent(persist)
        run(2)
        print(f"rep {rep} {name}: {run(10):.3f} ms/step", flush=True)
