"""Diagnostic: which decode mode disagrees with which (full arch, bf16, supplied uniforms)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import load_golden
from test_gpu_sampling import _sampler
from test_gpu_active_rows import _inputs, _first_eos
fx = load_golden("vaetf_full")
steps = 40
for n in (700, 1500):
    kw = _inputs(fx, n, 13, steps, seed=5)
    outs = {}
    for mode, extra in (("plainA", {}), ("plainB", {}), ("plain_nograph", dict(use_cuda_graph=False)),
                        ("skip", dict(skip_finished=True, compact_min_rows=10 ** 9, sync_every=5)),
                        ("skip_nograph", dict(skip_finished=True, compact_min_rows=10 ** 9, sync_every=5, use_cuda_graph=False)),
                        ("compact", dict(skip_finished=True, compact_min_rows=1, compact_every=3, compact_quantum=64))):
        s, _ = _sampler(fx, "bf16", algo="multinomial", max_strlen=steps + 1, **extra)
        with torch.no_grad():
            s.model.out.bias[3] += 2.0
        for rep in range(3):
            o = s._decode_cached(**kw).cpu()
            outs[f"{mode}{rep}"] = o
    ref = outs["plainA0"]
    e = _first_eos(ref)
    for k, o in outs.items():
        w = min(ref.size(1), o.size(1))
        cols = torch.arange(w)[None, :]
        live = cols <= e[:, None]
        diff = (ref[:, :w] != o[:, :w]) & live
        bad = diff.any(1)
        first = diff.float().argmax(1)[bad]
        print(f"n={n} {k:16s} width {o.size(1)} rows differing {int(bad.sum())} / {n}; first-diff step histogram {torch.bincount(first, minlength=w).tolist() if bad.any() else []}", flush=True)
