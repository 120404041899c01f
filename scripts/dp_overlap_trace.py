"""Evidence for the overlapped gradient exchange: under torchrun, runs cfg-4 data-parallel steps with
torch.profiler (CUPTI kernel records, device timestamps) and reports, for the last step of rank 0, how much of the NCCL
all-reduce kernels' time runs concurrently with this library's backward kernels, for the bucketed-overlap exchange and for
the single post-backward all-reduce.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dp_overlap_trace.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gct_plus_b200.Model import Cvaetf  # noqa: E402
from gct_plus_b200.Train.trainer1 import FusedTrainer  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
out = {}
for mode in ("overlap", "nccl"):
    torch.manual_seed(0)
    model = Cvaetf(32, 32, dropout=0.1, nconds=3, use_cond2lat=True, compute_dtype="bf16", **bench.ARCH).to(dev).train()
    tr = FusedTrainer(model, "pscavaetf", grad_exchange=mode)
    batch = bench.make_train_batch(512, 98, 3, 19, 1 + rank, dev=dev)
    for _ in range(3):
        tr.step(batch, 0.5)
    torch.cuda.synchronize()
    dist.barrier()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        tr.step(batch, 0.5)
        torch.cuda.synchronize()
    ev = [(e.name, e.time_range.start, e.time_range.end) for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    nccl = [(s, e) for n, s, e in ev if "nccl" in n.lower()]
    ours = [(n, s, e) for n, s, e in ev if "nccl" not in n.lower()]
    t0 = min(s for _, s, _ in ev)
    t1 = max(e for _, _, e in ev)
    nccl_total = sum(e - s for s, e in nccl)
    overlap = 0.0
    names = {}
    for s, e in nccl:
        for n, a, b in ours:
            o = min(e, b) - max(s, a)
            if o > 0:
                overlap += o
                k = n.split("(")[0][:60]
                names[k] = names.get(k, 0.0) + o
    last_ours = max(e for _, _, e in ours if True)
    bwd_end = max(e for n, _, e in ours if "adam" not in n)
    out[mode] = {"step_us": t1 - t0, "nccl_kernels": len(nccl), "nccl_busy_us": nccl_total,
                 "nccl_time_concurrent_with_compute_us": min(overlap, nccl_total), "fraction_concurrent": min(overlap, nccl_total) / max(nccl_total, 1e-9),
                 "nccl_first_start_us_into_step": (min(s for s, _ in nccl) - t0) if nccl else None,
                 "nccl_last_end_us_into_step": (max(e for _, e in nccl) - t0) if nccl else None,
                 "top_concurrent_kernels_us": dict(sorted(names.items(), key=lambda kv: -kv[1])[:6])}
    tr.xchg.close()
    del tr, model
    torch.cuda.empty_cache()
if rank == 0:
    print(json.dumps({"world": world, "workload": "cfg4 pscavaetf B=512/GPU S=98, one optimiser step, rank 0, CUPTI kernel timestamps", **out}))
dist.barrier()
dist.destroy_process_group()
