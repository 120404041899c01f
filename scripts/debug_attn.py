"""Error map of the tcgen05 attention forward / backward against torch fp32 (per gradient, batch, head, row block)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gct_plus_b200._lib as L  # noqa: E402

Lq, Lk, dense = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
B, H, d = 3, 2, 128
DEV = "cuda:0"
torch.manual_seed(0)
q = torch.randn(B, Lq, d, device=DEV).bfloat16()
k = torch.randn(B, Lk, d, device=DEV).bfloat16()
v = torch.randn(B, Lk, d, device=DEV).bfloat16()
if dense:
    mask = torch.tril(torch.ones(Lq, Lk, device=DEV, dtype=torch.bool)).expand(B, Lq, Lk).clone()
    mask[1, :, 3] = False
    mb, mr = Lq * Lk, Lk
else:
    lens = torch.tensor([Lk, max(1, Lk // 2), max(1, Lk - 3)], device=DEV)
    mask = (torch.arange(Lk, device=DEV)[None, :] < lens[:, None]).view(B, 1, Lk)
    mb, mr = Lk, 0
m8 = mask.to(torch.uint8).contiguous()
qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
qh = qf.view(B, Lq, H, 64).transpose(1, 2)
kh = kf.view(B, Lk, H, 64).transpose(1, 2)
vh = vf.view(B, Lk, H, 64).transpose(1, 2)
s = qh @ kh.transpose(-1, -2) / 8.0
s = s.masked_fill(~mask.view(B, 1, -1, Lk), -1e9)
pr = torch.softmax(s, -1)
ref = (pr @ vh).transpose(1, 2).reshape(B, Lq, d)
out = torch.empty(B, Lq, d, device=DEV, dtype=torch.bfloat16)
lse = torch.empty(B, H, Lq, device=DEV)
lib = L.lib()
L.check(lib.gct_attention_fwd(L.ptr(q), d, L.ptr(k), d, L.ptr(v), d, L.ptr(m8), mb, mr, L.ptr(out), d, L.ptr(lse), None,
                              B, H, Lq, Lk, 1, L.stream_ptr()))
torch.cuda.synchronize()
print("fwd max err", float((out.float() - ref).abs().max()))
dO = torch.randn(B, Lq, d, device=DEV).bfloat16()
ref.backward(dO.float())
dq, dk, dv = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
L.check(lib.gct_attention_bwd(L.ptr(q), d, L.ptr(k), d, L.ptr(v), d, L.ptr(m8), mb, mr, L.ptr(lse), L.ptr(out), d, L.ptr(dO), d,
                              L.ptr(dq), d, L.ptr(dk), d, L.ptr(dv), d, B, H, Lq, Lk, 1, L.stream_ptr()))
torch.cuda.synchronize()
for name, got, want in (("dq", dq, qf.grad), ("dk", dk, kf.grad), ("dv", dv, vf.grad)):
    e = (got.float() - want).abs()
    print(name, "max err", float(e.max()), "scale", float(want.abs().max()))
    for b in range(B):
        for h in range(H):
            eb = e[b, :, h * 64:(h + 1) * 64]
            rows = eb.max(dim=1).values
            bad = (rows > 0.05 * float(want.abs().max())).nonzero().flatten().tolist()
            cols = (eb.max(dim=0).values > 0.05 * float(want.abs().max())).nonzero().flatten().tolist()
            if bad:
                print(f"  b={b} h={h}: bad rows {bad[:12]}{'...' if len(bad) > 12 else ''} ({len(bad)}), bad cols {cols[:8]}... ({len(cols)})")
