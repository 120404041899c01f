"""Times the cfg-2 decode at the given batch sizes (device-resident inputs, CUDA graphs as in bench.py).
usage: time_decode.py [B ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

dev = torch.device("cuda:0")
if os.environ.get("GCT_DA_CFG") is not None:          # decode self-attention form: 0 tensor-core (default), -1 bulk-copy FMA kernels
    import gct_plus_b200._lib as L
    L.lib().gct_set_decode_attn_config(int(os.environ["GCT_DA_CFG"]))
for B in ([int(x) for x in sys.argv[1:]] or [512, 30000]):
    bench.BATCH = B
    s = bench.build_sampler(dev)
    toklen, zs = bench.sample_inputs(s, 1, seed=5, pinned=False)[0]
    Lz = zs.size(1)
    mask = (torch.arange(Lz).expand(B, 1, Lz) < torch.LongTensor(toklen).view(B, 1, 1)).to(dev)
    zs = zs.to(dev)
    ys0 = torch.full((B, 1), 2, dtype=torch.long, device=dev)
    for _ in range(3):
        s.decode(zs=zs, ys=ys0, src_mask=mask)
    reps = 10 if B <= 2048 else 4
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        s.decode(zs=zs, ys=ys0, src_mask=mask)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"B={B}: {ms:.2f} ms per call ({s.last_decode_steps} steps), {B / ms * 1e3:.0f} SMILES/s", flush=True)
    del s
    torch.cuda.empty_cache()
