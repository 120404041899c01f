"""Sweep of decode-attention kernel configurations (chunk / ring stages / rows per CTA) at fixed shapes."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gct_plus_b200._lib as L  # noqa: E402

dev = torch.device("cuda:0")
lib = L.lib()
d, H, Lmax, NL = 512, 8, 100, 6
for B in ([int(x) for x in sys.argv[1:]] or [512, 4096]):
    kc = torch.randn(NL, B, Lmax, d, device=dev).bfloat16()
    vc = torch.randn(NL, B, Lmax, d, device=dev).bfloat16()
    qkv = torch.randn(B, 3 * d, device=dev).bfloat16()
    out = torch.empty(B, d, device=dev, dtype=torch.bfloat16)
    valid = torch.ones(B, Lmax, device=dev, dtype=torch.uint8)
    for cfg in ([int(x) for x in os.environ.get("GCT_DA_CFGS", "").split(",") if x] or (2, 3, 4, 12, 831, 1621)):
        lib.gct_set_decode_attn_config(cfg)

        def launch(l, t):
            L.check(lib.gct_decode_attention(L.ptr(qkv), 3 * d, qkv[:, d:].data_ptr(), qkv[:, 2 * d:].data_ptr(), 3 * d, kc[l].data_ptr(),
                                             vc[l].data_ptr(), Lmax * d, d, t, L.ptr(valid), Lmax, L.ptr(out), d, B, H, 1, L.stream_ptr()))
        for t in (20, 49, 90):
            for l in range(NL):
                launch(l, t)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for rep in range(5):
                    for l in range(NL):
                        launch(l, t)
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / (5 * NL)
            by = B * (2 * t * d + 6 * d) * 2
            print(f"B={B} cfg={cfg} t={t}: {us:7.2f} us  {by / us / 1e3:7.1f} GB/s", flush=True)
