"""Deterministic weight fill shared by oracle/make_golden.py (which loads it into the
reference's own nn.Modules) and by the tests (which load it into the oracle and into the
CUDA implementation).  Keeps the committed fixtures small: weights are regenerated, only
inputs/outputs are stored."""
import math
import zlib

import torch


def fill_state_dict(shapes, seed: int = 0):
    """shapes: list of (key, shape) in state_dict order.  Returns {key: fp32 tensor}."""
    out = {}
    for key, shape in shapes:
        shape = tuple(shape)
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 31))
        if key.endswith("pe.pe"):
            out[key] = None          # buffer: filled by the caller with the PE table
            continue
        if key.endswith(".alpha"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif len(shape) >= 2:
            fan_out, fan_in = shape[0], shape[1]
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / (fan_in + fan_out))
        else:
            t = 0.05 * torch.randn(shape, generator=g)
        out[key] = t.float()
    return out
