"""A/B of the decode self-attention forms inside the whole cfg-2 decode (30 000 rows, device-resident inputs, interleaved):
tensor-core form (decode_attn_mma.cuh, gct_set_decode_attn_config(0 / 3 / 12)) vs the bulk-copy FMA kernels (-1)."""
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
s.use_cuda_graph = False          # the switch changes kernel choices: no stale graphs
toklen, zs = bench.sample_inputs(s, 1, seed=5, pinned=False)[0]
Lz = zs.size(1)
mask = (torch.arange(Lz).expand(B, 1, Lz) < torch.LongTensor(toklen).view(B, 1, 1)).to(dev)
zs = zs.to(dev)
ys0 = torch.full((B, 1), 2, dtype=torch.long, device=dev)
s.decode(zs=zs, ys=ys0, src_mask=mask)
for rep in range(3):
    for name, cfg in (("bulk-copy FMA kernels", -1), ("tensor-core form, 2 stages", 0), ("tensor-core form, one box per head", 12)):
        lib.gct_set_decode_attn_config(cfg)
        s.decode(zs=zs, ys=ys0, src_mask=mask)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s.decode(zs=zs, ys=ys0, src_mask=mask)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"rep {rep} B={B} {name:36s}: {ms:7.1f} ms per call, {B / ms * 1e3:7.0f} SMILES/s", flush=True)
lib.gct_set_decode_attn_config(0)
