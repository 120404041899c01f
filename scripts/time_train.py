"""usage: time_train.py [cfg3|cfg4] [steps] [batch].  Times the fused training step (CUDA events) with programmatic dependent
launch off / on (gct_set_pdl) -- and, run once per library build (GCT_B200_LIB=<path to an alternative libgct_b200.so>), gives an
A/B of two builds on the same box."""
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
case = bench.TRAIN_CASES[sys.argv[1] if len(sys.argv) > 1 else "cfg3"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
model = Cvaetf(32, 32, dropout=0.1, nconds=3, use_cond2lat=True, compute_dtype="bf16", **bench.ARCH).to(dev).train()
tr = FusedTrainer(model, case["mt"])
B = int(sys.argv[3]) if len(sys.argv) > 3 else 512
batch = bench.make_train_batch(B, case["S"], 3, case["sca"], 1, dev=dev)
for _ in range(5):
    tr.step(batch, 0.5)
for rep in range(3):
    for pdl in (0, 1):
        L.lib().gct_set_pdl(pdl)
        tr.step(batch, 0.5)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            tr.step(batch, 0.5)
        e1.record()
        torch.cuda.synchronize()
        print(f"{os.path.basename(L.LIB_PATH)} B={B} rep {rep} pdl={pdl}: {e0.elapsed_time(e1) / steps:.3f} ms/step", flush=True)
print("loss", tr.read_losses())
