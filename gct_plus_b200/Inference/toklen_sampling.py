"""Target-length sampler (drop-in for Inference/toklen_sampling.py): draws lengths from the
training-set histogram with a half-bin Gaussian jitter.  Host-side NumPy; consumes the global
NumPy RNG in the same order as the reference (one uniform, then one normal, per draw)."""
import numpy as np


def tokenlen_gen_from_data_distribution(data, size, nBins):
    counts, edges = np.histogram(data, bins=nBins)
    pdf = counts / np.sum(counts)
    width = np.diff(edges)[0]
    centres = edges[:-1] + 0.5 * width
    cdf = np.zeros_like(edges)
    cdf[1:] = np.cumsum(pdf)
    out = np.empty((size, 1))
    for k in range(size):
        a = np.random.uniform(0, 1)
        idx = np.argmax(cdf >= a) - 1
        out[k, 0] = centres[idx] + width * np.random.normal() / 2
    return out
