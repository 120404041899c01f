"""One latent-space cross-attention launch at decode shape (B rows, 8 heads, latent 128, ragged key counts) for ncu / timing."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device("cuda:0")
bench.BATCH = B
s = bench.build_sampler(dev)
s.use_cuda_graph = False
s.max_strlen = 6
inputs = bench.sample_inputs(s, 1, seed=5, pinned=False)
toklen, zs = inputs[0]
Lz = zs.size(1)
mask = (torch.arange(Lz).expand(B, 1, Lz) < torch.LongTensor(toklen).view(B, 1, 1)).to(dev)
ys0 = torch.full((B, 1), 2, dtype=torch.long, device=dev)
for _ in range(2):
    s.decode(zs=zs.to(dev), ys=ys0, src_mask=mask)
torch.cuda.synchronize()
print("ok: ran", s.last_decode_steps, "decode steps at B =", B)
