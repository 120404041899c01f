"""In-process A/B of the data-parallel gradient exchange on cfg 4 (torchrun, N >= 2): one all-reduce after the backward vs the
bucketed exchange overlapped with it, the latter with the persistent GEMMs held to fewer SMs.  CUDA events, max over ranks.
NCCL_MAX_CTAS (environment, read once per process by NCCL) bounds the CTAs of the all-reduce kernels."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import gct_plus_b200._lib as L  # noqa: E402
from gct_plus_b200.Model import Cvaetf  # noqa: E402
from gct_plus_b200.Train.trainer1 import FusedTrainer  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = L.lib()
torch.manual_seed(0)
model = Cvaetf(32, 32, dropout=0.1, nconds=3, use_cond2lat=True, compute_dtype="bf16", **bench.ARCH).to(dev).train()
batch = bench.make_train_batch(512, 98, 3, 19, 1 + rank, dev=dev)
cases = [("none (no exchange)", "none", 0), ("nccl after backward", "nccl", 0), ("overlap", "overlap", 0), ("overlap, 144 SMs", "overlap", 144),
         ("overlap, 140 SMs", "overlap", 140), ("overlap, 132 SMs", "overlap", 132), ("nccl after backward", "nccl", 0), ("overlap", "overlap", 0)]
for name, mode, budget in cases:
    tr = FusedTrainer(model, "pscavaetf", grad_exchange=mode)
    lib.gct_set_sm_budget(budget)
    for _ in range(4):
        tr.step(batch, 0.5)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        tr.step(batch, 0.5)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 20], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    lib.gct_set_sm_budget(0)
    if rank == 0:
        print(f"world {world} NCCL_MAX_CTAS={os.environ.get('NCCL_MAX_CTAS', '-')}  {name:28s} {float(t):8.3f} ms/step", flush=True)
    if tr.xchg is not None:
        tr.xchg.close()
    del tr
dist.barrier()
dist.destroy_process_group()
