"""Runs the decode attention kernel alone at a mid-sequence shape (target for `ncu --set full`)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gct_plus_b200._lib as L  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
t = int(sys.argv[2]) if len(sys.argv) > 2 else 49
dev = torch.device("cuda:0")
lib = L.lib()
d, H, Lmax, NL = 512, 8, 100, 3
kc = torch.randn(NL, B, Lmax, d, device=dev).bfloat16()
vc = torch.randn(NL, B, Lmax, d, device=dev).bfloat16()
qkv = torch.randn(B, 3 * d, device=dev).bfloat16()
out = torch.empty(B, d, device=dev, dtype=torch.bfloat16)
valid = torch.ones(B, Lmax, device=dev, dtype=torch.uint8)
for i in range(6):
    l = i % NL
    L.check(lib.gct_decode_attention(L.ptr(qkv), 3 * d, qkv[:, d:].data_ptr(), qkv[:, 2 * d:].data_ptr(), 3 * d, kc[l].data_ptr(),
                                     vc[l].data_ptr(), Lmax * d, d, t, L.ptr(valid), Lmax, L.ptr(out), d, B, H, 1, L.stream_ptr()))
torch.cuda.synchronize()
print("algorithmic bytes per launch:", B * (2 * t * d + 3 * d + 2 * d + d) * 2)
