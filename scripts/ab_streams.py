"""Decode throughput vs the number of concurrent row groups (Sampling.decode_streams) at small batches (cfg 2 shapes)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

dev = torch.device("cuda:0")
for B in (128, 512, 1024, 2048, 4096, 8192):
    bench.BATCH = B
    for G in (1, 2, 4, 8, 16):
        if B // G < 32:
            continue
        s = bench.build_sampler(dev)
        s.decode_streams = G
        toklen, zs = bench.sample_inputs(s, 1, seed=5, pinned=False)[0]
        Lz = zs.size(1)
        mask = (torch.arange(Lz).expand(B, 1, Lz) < torch.LongTensor(toklen).view(B, 1, 1)).to(dev)
        zs = zs.to(dev)
        ys0 = torch.full((B, 1), 2, dtype=torch.long, device=dev)
        for _ in range(3):
            s.decode(zs=zs, ys=ys0, src_mask=mask)
        reps = max(2, 4096 // B)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            s.decode(zs=zs, ys=ys0, src_mask=mask)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"B={B:5d} groups={G:2d}: {ms:8.2f} ms per call, {B / ms * 1e3:9.0f} SMILES/s", flush=True)
        del s
        torch.cuda.empty_cache()
