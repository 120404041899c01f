"""cfg-2 decode (30 000 rows, device-resident inputs, latent length padded to 56): latent-space cross-attention with CTAs of 3 warps
(five per SM, gct_set_zattn_config(0)) vs CTAs of 4 warps (three per SM at 14 KB per warp, config 5); and latent_bucket 64 for reference."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import gct_plus_b200._lib as L  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
bench.BATCH = B
lib = L.lib()
samplers = {bk: bench.build_sampler(dev, latent_bucket=bk) for bk in (8, 64)}
for s in samplers.values():
    s.use_cuda_graph = False
toklen, zs = bench.sample_inputs(samplers[8], 1, seed=5, pinned=False)[0]
Lz = zs.size(1)
mask = (torch.arange(Lz).expand(B, 1, Lz) < torch.LongTensor(toklen).view(B, 1, 1)).to(dev)
zs = zs.to(dev)
ys0 = torch.full((B, 1), 2, dtype=torch.long, device=dev)
for rep in range(3):
    for bk, cfg, name in ((8, 5, "56 keys, 4-warp CTAs"), (8, 0, "56 keys, 3-warp CTAs"), (64, 0, "64 keys, 4-warp CTAs")):
        lib.gct_set_zattn_config(cfg)
        s = samplers[bk]
        s.decode(zs=zs, ys=ys0, src_mask=mask)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s.decode(zs=zs, ys=ys0, src_mask=mask)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"rep {rep} B={B} {name}: {ms:7.1f} ms per call, {B / ms * 1e3:7.0f} SMILES/s", flush=True)
lib.gct_set_zattn_config(0)
