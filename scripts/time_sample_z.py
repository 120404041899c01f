"""Where the time of the seed-faithful latent draw goes (host MT19937 fill, H2D, device Box-Muller) next to torch.normal."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import gct_plus_b200._lib as L  # noqa: E402

dev = torch.device("cuda:0")
bench.BATCH = 30000
s = bench.build_sampler(dev)
lib = L.lib()
n, Lz, lat = 30000, 55, 128
numel = n * Lz * lat
torch.manual_seed(1)
for rep in range(3):
    t0 = time.perf_counter(); z = s.sample_z(Lz, n); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"sample_z (fast path) call {rep}: {(t1 - t0) * 1e3:.1f} ms")
raw = s._raw_host
state = torch.get_rng_state(); sn = state.numpy()
t0 = time.perf_counter()
L.check(lib.gct_mt19937_fill(sn[24:24 + 4992].view(np.uint64).ctypes.data, sn[8:12].view(np.int32).ctypes.data, sn[16:24].view(np.uint64).ctypes.data, raw.data_ptr(), numel))
t1 = time.perf_counter()
rd = raw[:numel].to(dev, non_blocking=True); torch.cuda.synchronize(); t2 = time.perf_counter()
zz = torch.empty(n, Lz, lat, device=dev)
L.check(lib.gct_normal_from_mt(L.ptr(rd), L.ptr(zz), numel, L.stream_ptr())); torch.cuda.synchronize(); t3 = time.perf_counter()
print(f"fill {1e3 * (t1 - t0):.1f} ms, H2D {1e3 * (t2 - t1):.1f} ms, Box-Muller {1e3 * (t3 - t2):.1f} ms")
for thr in (1, os.cpu_count()):
    torch.set_num_threads(thr)
    t0 = time.perf_counter(); zc = torch.normal(mean=0, std=1, size=(n, Lz, lat)); t1 = time.perf_counter()
    zd = zc.to(dev); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"torch.normal ({thr} threads): {1e3 * (t1 - t0):.1f} ms, pageable H2D {1e3 * (t2 - t1):.1f} ms")
for rep in range(2):
    np.random.seed(rep)
    t0 = time.perf_counter(); tl = s.sample_toklen(n); t1 = time.perf_counter()
    m = s._latent_mask(tl, n, int(max(tl))); t2 = time.perf_counter()
    print(f"sample_toklen {1e3 * (t1 - t0):.1f} ms, latent mask {1e3 * (t2 - t1):.1f} ms")
