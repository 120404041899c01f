"""A/B of library switches on the cfg-2 decode (device-resident inputs, one 30000-row call each, interleaved)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import gct_plus_b200._lib as L  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
dev = torch.device("cuda:0")
bench.BATCH = B
lib = L.lib()
s = bench.build_sampler(dev)
s.use_cuda_graph = False          # switches change kernel choices: no stale graphs
toklen, zs = bench.sample_inputs(s, 1, seed=5, pinned=False)[0]
Lz = zs.size(1)
mask = (torch.arange(Lz).expand(B, 1, Lz) < torch.LongTensor(toklen).view(B, 1, 1)).to(dev)
zs = zs.to(dev)
ys0 = torch.full((B, 1), 2, dtype=torch.long, device=dev)
s.decode(zs=zs, ys=ys0, src_mask=mask)
for rep in range(2):
    for name, pair in (("zattn rows from global memory, 3 CTAs/SM", 3), ("zattn rows staged in shared memory", 0)):
        lib.gct_set_zattn_config(pair)
        s.decode(zs=zs, ys=ys0, src_mask=mask)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s.decode(zs=zs, ys=ys0, src_mask=mask)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"rep {rep} {name}: {ms:.1f} ms per call, {B / ms * 1e3:.0f} SMILES/s", flush=True)
