import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gct_plus_b200._lib as L
Lq = Lk = int(sys.argv[1])
B, H, d = 2, 2, 128
DEV = "cuda:0"
lib = L.lib()
q = torch.zeros(B, Lq, d, device=DEV).bfloat16()
k = torch.randn(B, Lk, d, device=DEV).bfloat16()
v = torch.randn(B, Lk, d, device=DEV).bfloat16()
out = torch.empty(B, Lq, d, device=DEV, dtype=torch.bfloat16)
lse = torch.empty(B, H, Lq, device=DEV)
L.check(lib.gct_attention_fwd(L.ptr(q), d, L.ptr(k), d, L.ptr(v), d, None, 0, 0, L.ptr(out), d, L.ptr(lse), None, B, H, Lq, Lk, 1, L.stream_ptr()))
torch.cuda.synchronize()
print("lse", lse[0, 0, :3].tolist(), "expect", float(torch.log(torch.tensor(float(Lk)))))
for r0 in range(0, Lq, 16):
    dO = torch.zeros(B, Lq, d, device=DEV).bfloat16()
    dO[:, r0:r0 + 16, :] = 1.0
    dq, dk, dv = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
    L.check(lib.gct_attention_bwd(L.ptr(q), d, L.ptr(k), d, L.ptr(v), d, None, 0, 0, L.ptr(lse), L.ptr(out), d, L.ptr(dO), d,
                                  L.ptr(dq), d, L.ptr(dk), d, L.ptr(dv), d, B, H, Lq, Lk, 1, L.stream_ptr()))
    torch.cuda.synchronize()
    want = min(16, Lq - r0) / Lk
    g = dv[0, :, :64].float()
    print(f"query block {r0}: dv want {want:.4f}; per key-block mean:", [round(float(g[j:j + 16].mean()), 4) for j in range(0, Lk, 16)],
          "col-block mean:", [round(float(g[:, c:c + 16].mean()), 4) for c in range(0, 64, 16)])
