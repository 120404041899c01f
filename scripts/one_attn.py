"""One cfg-3-shaped attention forward + backward (B=512, H=8, L given, fused qkv pitch 1536) for ncu / timing."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gct_plus_b200._lib as L  # noqa: E402

Lq = int(sys.argv[1]) if len(sys.argv) > 1 else 81
dense = int(sys.argv[2]) if len(sys.argv) > 2 else 0
B, H, d = 512, 8, 512
dev = torch.device("cuda:0")
lib = L.lib()
qkv = torch.randn(B * Lq, 3 * d, device=dev).bfloat16()
out = torch.empty(B * Lq, d, device=dev, dtype=torch.bfloat16)
dO = torch.randn(B * Lq, d, device=dev).bfloat16()
dqkv = torch.empty_like(qkv)
lse = torch.empty(B, H, Lq, device=dev)
if dense:
    m8 = torch.tril(torch.ones(Lq, Lq, device=dev, dtype=torch.uint8)).expand(B, Lq, Lq).contiguous()
    mb, mr = Lq * Lq, Lq
else:
    m8 = torch.ones(B, Lq, device=dev, dtype=torch.uint8)
    m8[:, Lq - 5:] = 0
    mb, mr = Lq, 0
q, k, v = qkv, qkv[:, d:], qkv[:, 2 * d:]
dq, dk, dv = dqkv, dqkv[:, d:], dqkv[:, 2 * d:]
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for it in range(4):
    if it == 3:
        ev[0].record()
    L.check(lib.gct_attention_fwd(L.ptr(q), 3 * d, L.ptr(k), 3 * d, L.ptr(v), 3 * d, L.ptr(m8), mb, mr, L.ptr(out), d, L.ptr(lse), None,
                                  B, H, Lq, Lq, 1, L.stream_ptr()))
    if it == 3:
        ev[1].record()
    L.check(lib.gct_attention_bwd(L.ptr(q), 3 * d, L.ptr(k), 3 * d, L.ptr(v), 3 * d, L.ptr(m8), mb, mr, L.ptr(lse), L.ptr(out), d,
                                  L.ptr(dO), d, L.ptr(dq), 3 * d, L.ptr(dk), 3 * d, L.ptr(dv), 3 * d, B, H, Lq, Lq, 1, L.stream_ptr()))
ev[2].record()
torch.cuda.synchronize()
print(f"attention L={Lq} dense={dense}: fwd {ev[0].elapsed_time(ev[1]) * 1e3:.1f} us, bwd {ev[1].elapsed_time(ev[2]) * 1e3:.1f} us")
