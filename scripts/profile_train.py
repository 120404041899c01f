"""One cfg-3 optimiser step (pvaetf, B=512, S=78, T=79, bf16) for ncu launch lists."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gct_plus_b200.Model import Cvaetf  # noqa: E402
from gct_plus_b200.Train.trainer1 import FusedTrainer  # noqa: E402

dev = torch.device("cuda:0")
if os.environ.get("GCT_FFN_PREACT") is not None:
    import gct_plus_b200._lib as L
    L.lib().gct_set_ffn_saved_activation(int(os.environ["GCT_FFN_PREACT"]))
torch.manual_seed(0)
B = int(os.environ.get("GCT_PROFILE_B", "512"))
model = Cvaetf(32, 32, dropout=0.1, nconds=3, use_cond2lat=True, compute_dtype="bf16", **bench.ARCH).to(dev).train()
tr = FusedTrainer(model, "pvaetf")
batch = bench.make_train_batch(B, 78, 3, 0, 1, dev=dev)
for _ in range(2):
    tr.step(batch, 0.5)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.step(batch, 0.5)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one train step; loss", tr.read_losses())
