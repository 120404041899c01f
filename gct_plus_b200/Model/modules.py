"""Parameter containers and mask builders mirroring Model/modules.py of the reference.

The containers only *hold* parameters under the reference's attribute names (so that
``state_dict()`` keys, shapes and the xavier-init RNG order are identical); the
arithmetic lives in the sm_100a kernels reached through ``engine.py``.  Calling a container
directly runs the corresponding single kernel (used by the op-level parity tests).
"""
from __future__ import annotations

import copy
import math

import torch
import torch.nn as nn

from .. import _lib as L

_PE_CACHE = {}


def positional_table(d_model: int, max_seq_len: int = 200) -> torch.Tensor:
    """Table of Model/modules.py:122-130, evaluated with the same python-float arithmetic so the
    registered buffer is bit-identical: sin(pos/10000^(2i/d)) at even i, cos(pos/10000^(2(i+1)/d))."""
    key = (d_model, max_seq_len)
    if key not in _PE_CACHE:
        rows = []
        inv_sin = [10000 ** ((2 * i) / d_model) for i in range(0, d_model, 2)]
        inv_cos = [10000 ** ((2 * (i + 1)) / d_model) for i in range(0, d_model, 2)]
        for pos in range(max_seq_len):
            row = [0.0] * d_model
            for k in range(d_model // 2):
                row[2 * k] = math.sin(pos / inv_sin[k])
                row[2 * k + 1] = math.cos(pos / inv_cos[k])
            rows.append(row)
        _PE_CACHE[key] = torch.tensor(rows, dtype=torch.float64).float().unsqueeze(0)
    return _PE_CACHE[key].clone()


def get_clones(layer, N):
    return nn.ModuleList([copy.deepcopy(layer) for _ in range(N)])


class Norm(nn.Module):
    """alpha*(x-mean)/(std_unbiased+eps)+bias  (reference Model/modules.py:80-95)."""

    def __init__(self, d_model, eps=1e-6):
        super().__init__()
        self.size = d_model
        self.alpha = nn.Parameter(torch.ones(self.size))
        self.bias = nn.Parameter(torch.zeros(self.size))
        self.eps = eps

    def forward(self, x):
        L.require_cuda(x, "x")
        x32 = x.float().contiguous()
        y = torch.empty_like(x32)
        rows = x32.numel() // self.size
        L.check(L.lib().gct_norm_fwd(L.ptr(x32), L.ptr(self.alpha.data.float()), L.ptr(self.bias.data.float()), L.ptr(y),
                                     None, rows, self.size, L.DTYPE_F32, L.stream_ptr()), "gct_norm_fwd")
        return y


class Embeddings(nn.Module):
    def __init__(self, d_model, vocab):
        super().__init__()
        self.embed = nn.Embedding(vocab, d_model)
        self.d_model = d_model


class PositionalEncoding(nn.Module):
    def __init__(self, d_model, max_seq_len=200, dropout=0.1):
        super().__init__()
        self.d_model = d_model
        self.dropout = nn.Dropout(dropout)
        self.register_buffer("pe", positional_table(d_model, max_seq_len))


# ---------------------------------------------------------------------------------------------
# masks (reference Model/modules.py:10-66), built on the device by gct_src_mask / gct_trg_mask
# ---------------------------------------------------------------------------------------------
def get_cond_mask(conditions):
    return torch.ones_like(torch.unsqueeze(conditions, -2), dtype=torch.bool)


def get_src_mask(src, pad_idx, conditions=None):
    """(B,1,nc+S) bool: True where the key may be attended (cond columns always True)."""
    L.require_cuda(src, "src")
    src = src.contiguous()
    B, S = src.shape
    nc = 0 if conditions is None else conditions.size(-1)
    out = torch.empty((B, 1, nc + S), dtype=torch.uint8, device=src.device)
    L.check(L.lib().gct_src_mask(L.ptr(src), B, S, nc, int(pad_idx), L.ptr(out), L.stream_ptr()), "gct_src_mask")
    return out.view(torch.bool)


def get_trg_mask(target, pad_id, use_cond2dec, conditions=None):
    """(B,L,L) padding & no-peak mask, L = T (+nc with cond2dec).  The reference returns int64
    {0,1} (bool & bool*pad_idx with <pad> == 1); the truth values are the same, here as bool."""
    L.require_cuda(target, "target")
    target = target.contiguous()
    B, T = target.shape
    nc = conditions.size(-1) if (use_cond2dec and conditions is not None) else 0
    W = nc + T
    out = torch.empty((B, W, W), dtype=torch.uint8, device=target.device)
    L.check(L.lib().gct_trg_mask(L.ptr(target), B, T, nc, int(pad_id), L.ptr(out), L.stream_ptr()), "gct_trg_mask")
    return out.view(torch.bool)


def get_masks(source, target, conditions, pad_idx, use_cond2dec=True):
    return get_src_mask(source, pad_idx, conditions), get_trg_mask(target, pad_idx, use_cond2dec, conditions)


def mask_to_bytes(mask: torch.Tensor, shape) -> torch.Tensor:
    """Any caller-built mask (bool / uint8 / int64, broadcastable to `shape`) -> contiguous uint8."""
    L.require_cuda(mask, "mask")
    if tuple(mask.shape) != tuple(shape):
        mask = mask.expand(shape)
    mask = mask.contiguous()
    if mask.dtype in (torch.bool, torch.uint8):
        return mask.view(torch.uint8)
    if mask.dtype not in (torch.int32, torch.int64):
        raise L.GctError(f"unsupported mask dtype {mask.dtype}")
    out = torch.empty(shape, dtype=torch.uint8, device=mask.device)
    L.check(L.lib().gct_mask_cast(L.ptr(mask), mask.element_size(), mask.numel(), L.ptr(out), L.stream_ptr()), "gct_mask_cast")
    return out
