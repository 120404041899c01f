"""Per-CTA phase timeline of the tcgen05 attention kernels at the cfg-3 training shape (gct_set_attention_trace).
Phases: 0 start, 1 TMEM allocated (+ mask bits / lse requested in the backward), 2 operand tiles landed, 3 first MMAs done,
4 softmax pass done, 5 last MMAs done, 6 outputs stored.  Prints mean phase durations and the idle gap between successive CTAs
of one SM
(for the persistent kernels: between successive tiles of the SM's CTAs)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gct_plus_b200._lib as L  # noqa: E402

dev = torch.device("cuda:0")
lib = L.lib()
B, H, Lq, Lk = 512, 8, int(sys.argv[1]) if len(sys.argv) > 1 else 81, int(sys.argv[2]) if len(sys.argv) > 2 else 81
d = H * 64
torch.manual_seed(0)
q, k, v, do = (torch.randn(B, Lq if i in (0, 3) else Lk, d, device=dev).bfloat16() for i in range(4))
mask = torch.ones(B, 1, Lk, dtype=torch.uint8, device=dev)
o = torch.empty(B, Lq, d, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, H, Lq, device=dev)
dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
trace = torch.zeros(B * H, 16, dtype=torch.int64, device=dev)


def fwd():
    L.check(lib.gct_attention_fwd(L.ptr(q), d, L.ptr(k), d, L.ptr(v), d, L.ptr(mask), Lk, 0, L.ptr(o), d, L.ptr(lse), None, B, H, Lq, Lk, 1,
                                  L.stream_ptr()))


def bwd():
    L.check(lib.gct_attention_bwd(L.ptr(q), d, L.ptr(k), d, L.ptr(v), d, L.ptr(mask), Lk, 0, L.ptr(lse), L.ptr(o), d, L.ptr(do), d, L.ptr(dq), d,
                                  L.ptr(dk), d, L.ptr(dv), d, B, H, Lq, Lk, 1, L.stream_ptr()))


def report(name, fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 200
    lib.gct_set_attention_trace(L.ptr(trace))
    fn()
    torch.cuda.synchronize()
    lib.gct_set_attention_trace(None)
    tr = trace.cpu().numpy()
    g, c, sm = tr[:, :7].astype(np.float64), tr[:, 8:15].astype(np.float64), tr[:, 7]
    marked = [i for i in range(7) if g[:, i].any()]          # the persistent kernels do not mark phase 1
    print(f"{name}: {us:.1f} us/launch; kernel span by globaltimer {(g[:, 6].max() - g[:, 0].min()) / 1e3:.1f} us")
    print("  mean phase durations (SM clocks): " + "  ".join(f"{a}->{b}: {(c[:, b] - c[:, a]).mean():7.0f}" for a, b in zip(marked, marked[1:]))
          + f"   total {(c[:, 6] - c[:, 0]).mean():.0f}")
    print("  mean phase durations (ns):        " + "  ".join(f"{a}->{b}: {(g[:, b] - g[:, a]).mean():7.0f}" for a, b in zip(marked, marked[1:]))
          + f"   total {(g[:, 6] - g[:, 0]).mean():.0f}")
    # per SM: sort CTAs by start, pair each with the co-resident slots -> idle between an end and the next start on this SM
    gaps, conc = [], []
    for s in np.unique(sm):
        idx = np.where(sm == s)[0]
        st, en = np.sort(g[idx, 0]), np.sort(g[idx, 6])
        busy = (g[idx, 6] - g[idx, 0]).sum()
        conc.append(busy / (en[-1] - st[0]))
        # k-th start after the first wave follows the (k - wave)-th end
        wave = int(np.sum(st < en[0]))
        if len(st) > wave:
            gaps.append((st[wave:] - en[:len(st) - wave]).mean())
    print(f"  CTAs per SM {len(sm) / len(np.unique(sm)):.1f}; mean resident CTAs per SM {np.mean(conc):.2f}; mean end -> next start on the SM "
          f"{np.mean(gaps):.0f} ns")


for persistent in (0, 1):
    lib.gct_set_attention_persistent(3 if persistent else 0)
    trace.zero_()
    report(f"attention forward (no dropout), persistent={persistent}", fwd)
    trace.zero_()
    report(f"attention backward (no dropout, no bias sums), persistent={persistent}", bwd)

# the last attention launch of a training step = encoder layer 0's self-attention backward, with dropout and bias-gradient sums
import bench  # noqa: E402
from gct_plus_b200.Model import Cvaetf  # noqa: E402
from gct_plus_b200.Train.trainer1 import FusedTrainer  # noqa: E402

model = Cvaetf(32, 32, dropout=0.1, nconds=3, use_cond2lat=True, compute_dtype="bf16", **bench.ARCH).to(dev).train()
tr = FusedTrainer(model, "pvaetf")
batch = bench.make_train_batch(512, 78, 3, 0, 1, dev=dev)
for persistent in (0, 1, 0, 1):
    lib.gct_set_attention_persistent(3 if persistent else 0)
    trace.zero_()
    report(f"training step (trace = its last attention backward), persistent={persistent}", lambda: tr.step(batch, 0.5))
lib.gct_set_attention_persistent(2)
