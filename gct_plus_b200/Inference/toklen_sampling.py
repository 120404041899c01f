"""Target-length sampler (drop-in for Inference/toklen_sampling.py): draws lengths from the
training-set histogram with a half-bin Gaussian jitter.  The per-draw loop of the reference (one
np.random.uniform, then one np.random.normal, per draw) runs in the library's host code
(gct_toklen_draw) on NumPy's own global MT19937 state, so a seeded run draws exactly the reference's
lengths and leaves np.random in exactly the reference's state -- 30 000 draws take ~1 ms instead of
~0.2 s of Python."""
import ctypes as C

import numpy as np


def _python_loop(cdf, centres, width, size):
    """The reference's loop, kept as the fallback for a non-MT19937 global bit generator."""
    out = np.empty((size, 1))
    for k in range(size):
        a = np.random.uniform(0, 1)
        idx = np.argmax(cdf >= a) - 1
        out[k, 0] = centres[idx] + width * np.random.normal() / 2
    return out


def tokenlen_gen_from_data_distribution(data, size, nBins):
    counts, edges = np.histogram(data, bins=nBins)
    pdf = counts / np.sum(counts)
    width = np.diff(edges)[0]
    centres = edges[:-1] + 0.5 * width
    cdf = np.zeros_like(edges)
    cdf[1:] = np.cumsum(pdf)
    state = np.random.get_state()
    if state[0] != 'MT19937' or size == 0:
        return _python_loop(cdf, centres, width, size)
    from .. import _lib as L
    key = np.ascontiguousarray(state[1], dtype=np.uint32).copy()
    pos, has_gauss, cached = C.c_int32(int(state[2])), C.c_int32(int(state[3])), C.c_double(float(state[4]))
    cdf_c = np.ascontiguousarray(cdf, dtype=np.float64)
    cen_c = np.ascontiguousarray(centres, dtype=np.float64)
    out = np.empty((size, 1), dtype=np.float64)
    L.check(L.lib().gct_toklen_draw(key.ctypes.data, C.addressof(pos), C.addressof(has_gauss), C.addressof(cached),
                                    cdf_c.ctypes.data, len(cdf_c), cen_c.ctypes.data, float(width), int(size), out.ctypes.data),
            "gct_toklen_draw")
    np.random.set_state(('MT19937', key, int(pos.value), int(has_gauss.value), float(cached.value)))
    return out
