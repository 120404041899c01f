"""One mid-sequence window of the cfg-2 decode (steps 48..51 of a 512-draw batch) for ncu launch lists."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import gct_plus_b200._lib as L  # noqa: E402

dev = torch.device("cuda:0")
bench.BATCH = int(os.environ.get("GCT_PROFILE_B", "512"))
s = bench.build_sampler(dev)
s.use_cuda_graph = False
s.decode_streams = 1          # one row group: the window below drives the library directly
inputs = bench.sample_inputs(s, 2, seed=5, pinned=False)
toklen, zs = inputs[0]
NB = zs.size(0)
Lz = zs.size(1)
mask = (torch.arange(Lz).expand(NB, 1, Lz) < torch.LongTensor(toklen).view(NB, 1, 1)).to(dev)
ys0 = torch.full((NB, 1), 2, dtype=torch.long, device=dev)
s.decode(zs=zs.to(dev), ys=ys0, src_mask=mask)          # warm-up (all attributes set, descriptors cached)
torch.cuda.synchronize()
# second batch: run steps manually so that only a window is profiled
lib, model = L.lib(), s.model
cfg = model._cfg()
st = next(iter(s._static.values()))
ws = model._ws.get(('decode', 0), 0, dev)
w = model._weights()
dec = L.GctDecode(B=NB, Lz=st['zs'].size(1), max_len=st['ys'].size(1), prefix_len=1, greedy=0, eos_id=3, seed=0,
                  zs=st['zs'].data_ptr(), src_mask=st['mask'].data_ptr(), dconds=None, uniforms=st['uni'][0].data_ptr(),
                  ys=st['ys'].data_ptr(), status=st['status'].data_ptr())
L.check(lib.gct_decode_begin(C.byref(cfg), C.byref(w), C.byref(dec), L.ptr(ws), ws.numel(), L.stream_ptr()))
L.check(lib.gct_decode_steps(C.byref(cfg), C.byref(w), C.byref(dec), 0, 48, L.ptr(ws), ws.numel(), L.stream_ptr()))
torch.cuda.synchronize()
torch.cuda.profiler.start()
L.check(lib.gct_decode_steps(C.byref(cfg), C.byref(w), C.byref(dec), 48, 52, L.ptr(ws), ws.numel(), L.stream_ptr()))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled 4 decode steps (48..51), launches per step:", lib.gct_decode_launches_per_step(C.byref(cfg)))
