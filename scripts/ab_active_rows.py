"""A/B of the active-row decode (sampler kwarg skip_finished) on the cfg-2 workload: one B-row call per setting, device-resident
inputs, interleaved.  usage: ab_active_rows.py [B] [eos_bias]   (eos_bias is added to the <eos> logit: 0 = the random-init model)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
bias = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
dev = torch.device("cuda:0")
bench.BATCH = B
s = bench.build_sampler(dev)
if bias:
    with torch.no_grad():
        s.model.out.bias[3] += bias
toklen, zs = bench.sample_inputs(s, 1, seed=5, pinned=False)[0]
Lz = zs.size(1)
mask = (torch.arange(Lz).expand(B, 1, Lz) < torch.LongTensor(toklen).view(B, 1, 1)).to(dev)
zs = zs.to(dev)
ys0 = torch.full((B, 1), 2, dtype=torch.long, device=dev)
settings = [("plain (every row, every step)", dict(skip_finished=False)),
            ("skip only (graphs, no gather)", dict(skip_finished=True, compact_min_rows=10 ** 9)),
            ("gather every 4 steps", dict(skip_finished=True, compact_min_rows=1, compact_every=4)),
            ("gather every 8 steps", dict(skip_finished=True, compact_min_rows=1, compact_every=8)),
            ("gather every 16 steps", dict(skip_finished=True, compact_min_rows=1, compact_every=16))]
for name, kw in settings:                 # warm-up: static buffers, graphs
    for k, v in kw.items():
        setattr(s, k, v)
    for _ in range(3):
        s.decode(zs=zs, ys=ys0, src_mask=mask)
for rep in range(2):
    for name, kw in settings:
        for k, v in kw.items():
            setattr(s, k, v)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s.decode(zs=zs, ys=ys0, src_mask=mask)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        frac = s.last_row_steps / max(1, s.last_steps_executed * B)
        print(f"rep {rep} B={B} eos_bias={bias} {name:32s}: {ms:8.1f} ms per call, {B / ms * 1e3:8.0f} SMILES/s, steps {s.last_decode_steps}, "
              f"row-steps {frac:.3f}", flush=True)
