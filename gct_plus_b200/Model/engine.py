"""Host-side engine shared by Vaetf and Cvaetf: flat parameter storage, the ctypes calls into
libgct_b200.so, and the autograd bridge.

Parameters stay ordinary ``nn.Parameter`` objects under the reference's names, but their storage
is one flat fp32 buffer in *kernel layout* (q/k/v weights adjacent so self-attention runs one
N=3d GEMM, mu/log_var heads adjacent, ...).  A bf16 shadow of that buffer feeds the tcgen05 GEMMs.
Backward is hand-written in CUDA (gct_backward): one autograd.Function wraps the whole network, so
the reference's training loop (loss.backward(); optimizer.step()) works unchanged, while
``Train.trainer1.FusedTrainer`` bypasses autograd entirely.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref

import numpy as np
import torch
import torch.nn as nn

from .. import _lib as L

_ALIGN = 64


def _default_dtype() -> str:
    return os.environ.get("GCT_B200_DTYPE", "bf16")


class _Workspace:
    """Grow-only device scratch buffers keyed by purpose."""

    def __init__(self):
        self.bufs = {}

    def get(self, key, nbytes, device):
        t = self.bufs.get(key)
        if t is None or t.numel() < nbytes or t.device != device:
            t = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
            self.bufs[key] = t
        return t


class TransformerVAE(nn.Module):
    """Base of Vaetf / Cvaetf.  Subclasses build the reference's module tree, then call _finalize()."""

    # ------------------------------------------------------------------ construction
    def _finalize(self, compute_dtype=None):
        self.compute_dtype = compute_dtype or _default_dtype()
        assert self.compute_dtype in ("fp32", "bf16")
        self._flat = None
        self._shadow = None
        self._shadow_fresh = False
        self._ws = _Workspace()
        self._graph_ws = []
        self.pad_id = 1            # <pad> of the torchtext vocab; set by get_model / samplers
        self._step_seed = None
        self._flatten()

    def _slot_entries(self):
        """[(slot, [(owner_module, attr_name, is_buffer), ...])]: tensors stored back-to-back per slot."""
        e, d = self.encoder, self.decoder
        N = len(e.layers)
        G = L.NUM_GLOBAL_SLOTS
        heads = self.sampler if hasattr(self, "sampler") else e
        ent = [(0, [(e.embed_sentence.embed, "weight", False)]),
               (3, [(e.norm, "alpha", False)]), (4, [(e.norm, "bias", False)]),
               (5, [(heads.fc_mu, "weight", False), (heads.fc_log_var, "weight", False)]),
               (6, [(heads.fc_mu, "bias", False), (heads.fc_log_var, "bias", False)]),
               (7, [(d.embed.embed, "weight", False)]),
               (12, [(d.fc_z, "weight", False)]), (13, [(d.fc_z, "bias", False)]),
               (14, [(d.norm, "alpha", False)]), (15, [(d.norm, "bias", False)]),
               (16, [(self.out, "weight", False)]), (17, [(self.out, "bias", False)]),
               (20, [(e.pe, "pe", True)]), (21, [(d.pe, "pe", True)])]
        if self.nconds > 0 and hasattr(e, "embed_cond2enc"):
            ent += [(1, [(e.embed_cond2enc, "weight", False)]), (2, [(e.embed_cond2enc, "bias", False)])]
        if hasattr(d, "embed_cond2dec"):
            ent += [(8, [(d.embed_cond2dec, "weight", False)]), (9, [(d.embed_cond2dec, "bias", False)])]
        if hasattr(d, "embed_cond2lat"):
            ent += [(10, [(d.embed_cond2lat, "weight", False)]), (11, [(d.embed_cond2lat, "bias", False)])]
        if hasattr(self, "prop_fc"):
            ent += [(18, [(self.prop_fc, "weight", False)]), (19, [(self.prop_fc, "bias", False)])]

        def mha(m):
            return ([(m.q_linear, "weight", False), (m.k_linear, "weight", False), (m.v_linear, "weight", False)],
                    [(m.q_linear, "bias", False), (m.k_linear, "bias", False), (m.v_linear, "bias", False)])

        for l, lay in enumerate(e.layers):
            b = G + l * L.ENC_LAYER_SLOTS
            qw, qb = mha(lay.attn)
            ent += [(b + 0, [(lay.norm_1, "alpha", False)]), (b + 1, [(lay.norm_1, "bias", False)]), (b + 2, qw), (b + 3, qb),
                    (b + 4, [(lay.attn.out, "weight", False)]), (b + 5, [(lay.attn.out, "bias", False)]),
                    (b + 6, [(lay.norm_2, "alpha", False)]), (b + 7, [(lay.norm_2, "bias", False)]),
                    (b + 8, [(lay.ff.linear_1, "weight", False)]), (b + 9, [(lay.ff.linear_1, "bias", False)]),
                    (b + 10, [(lay.ff.linear_2, "weight", False)]), (b + 11, [(lay.ff.linear_2, "bias", False)])]
        for l, lay in enumerate(d.layers):
            b = G + N * L.ENC_LAYER_SLOTS + l * L.DEC_LAYER_SLOTS
            qw, qb = mha(lay.attn_1)
            a2 = lay.attn_2
            ent += [(b + 0, [(lay.norm_1, "alpha", False)]), (b + 1, [(lay.norm_1, "bias", False)]), (b + 2, qw), (b + 3, qb),
                    (b + 4, [(lay.attn_1.out, "weight", False)]), (b + 5, [(lay.attn_1.out, "bias", False)]),
                    (b + 6, [(lay.norm_2, "alpha", False)]), (b + 7, [(lay.norm_2, "bias", False)]),
                    (b + 8, [(a2.q_linear, "weight", False)]), (b + 9, [(a2.q_linear, "bias", False)]),
                    (b + 10, [(a2.k_linear, "weight", False), (a2.v_linear, "weight", False)]),
                    (b + 11, [(a2.k_linear, "bias", False), (a2.v_linear, "bias", False)]),
                    (b + 12, [(a2.out, "weight", False)]), (b + 13, [(a2.out, "bias", False)]),
                    (b + 14, [(lay.norm_3, "alpha", False)]), (b + 15, [(lay.norm_3, "bias", False)]),
                    (b + 16, [(lay.ff.linear_1, "weight", False)]), (b + 17, [(lay.ff.linear_1, "bias", False)]),
                    (b + 18, [(lay.ff.linear_2, "weight", False)]), (b + 19, [(lay.ff.linear_2, "bias", False)])]
        return ent

    def _flatten(self):
        """(Re)builds the flat fp32 buffer on the parameters' current device and re-points every
        parameter / PE buffer at its slice.  Called after construction and after .to()/.cuda()."""
        ent = self._slot_entries()
        nslots = L.NUM_GLOBAL_SLOTS + len(self.encoder.layers) * (L.ENC_LAYER_SLOTS + L.DEC_LAYER_SLOTS)
        offsets = np.full(nslots, -1, dtype=np.int64)
        placed, cur, plan = set(), 0, []
        for slot, members in ent:
            offsets[slot] = cur
            for mod, name, is_buf in members:
                t = mod._buffers[name] if is_buf else mod._parameters[name]
                plan.append((t, cur, mod, name, is_buf))
                placed.add(id(t))
                cur += t.numel()
            cur = (cur + _ALIGN - 1) // _ALIGN * _ALIGN
        for p in self.parameters():          # parameters the kernels never read (e.g. vaetf's encoder.fc_mu)
            if id(p) not in placed:
                plan.append((p, cur, None, None, False))
                placed.add(id(p))
                cur = (cur + p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        device = next(self.parameters()).device
        flat = torch.zeros(cur, dtype=torch.float32, device=device)
        views = {}
        for t, off, mod, name, is_buf in plan:
            if t.dtype != torch.float32:
                raise L.GctError("gct_plus_b200 keeps fp32 master parameters; use compute_dtype for bf16 math")
            view = flat[off:off + t.numel()].view(t.shape)
            view.copy_(t.data)
            if is_buf:
                mod._buffers[name] = view
            else:
                t.data = view
                views[id(t)] = (off, t.numel(), tuple(t.shape))
        self._flat = flat
        self._offsets = offsets
        self._shadow = None
        self._shadow_fresh = False
        self._param_list = list(self.parameters())
        # (offset, numel, shape) per parameter, in model.parameters() order (index-keyed: survives deepcopy / pickling)
        self._grad_views = [views[id(p)] for p in self._param_list]

    def _apply(self, fn, recurse=True):
        super()._apply(fn, recurse)
        if getattr(self, "_flat", None) is not None:
            self._flatten()
        return self

    def set_compute_dtype(self, dtype: str):
        assert dtype in ("fp32", "bf16")
        self.compute_dtype = dtype
        return self

    # ------------------------------------------------------------------ C structs
    def _cfg(self, dropout=None) -> L.GctConfig:
        e = self.encoder
        return L.GctConfig(src_vocab=e.embed_sentence.embed.num_embeddings, trg_vocab=self.out.out_features,
                           n_layers=len(e.layers), d_model=self.out.in_features, d_ff=e.layers[0].ff.linear_1.out_features,
                           heads=e.layers[0].attn.h, latent_dim=self.decoder.fc_z.in_features, nconds=int(self.nconds),
                           use_cond2dec=int(bool(self.use_cond2dec)), use_cond2lat=int(bool(self.use_cond2lat)),
                           dtype=L.DTYPE_BF16 if self.compute_dtype == "bf16" else L.DTYPE_F32, pad_id=int(self.pad_id),
                           dropout=float(e.pe.dropout.p if dropout is None else dropout))

    def _aliased(self) -> bool:
        """True while every parameter still lives at its slot of the flat buffer (``p.data = other`` breaks that)."""
        base = self._flat.data_ptr()
        return all(p.data_ptr() == base + 4 * off for p, (off, _, _) in zip(self._param_list, self._grad_views))

    def sync_weights(self):
        """Re-establishes the kernel view of the parameters after out-of-band edits: re-flattens if a parameter was
        re-pointed (``p.data = t``, deepcopy, unpickling) and re-casts the bf16 operand copy.  Called automatically
        before every kernel call; public so that callers can force it (e.g. after writing into ``model._flat``)."""
        if len(self._param_list) != len(self._grad_views) or not self._aliased():
            self._flatten()
        self._shadow_fresh = False
        return self

    def _weights(self, grads=None, trust_shadow=False) -> L.GctWeights:
        """Pointer struct for the library.  The bf16 operand copy is re-cast from the fp32 masters on every call (44 M
        elements, ~50 us) -- autograd version counters miss writes through ``.data`` -- unless `trust_shadow`: the
        fused trainer's Adam kernel writes master and shadow in the same pass and vouches for it."""
        if not self._flat.is_cuda:
            raise L.GctError("model parameters are on the CPU: move the model to a CUDA device (no CPU path)")
        if not self._aliased():
            self._flatten()
        shadow = None
        if self.compute_dtype == "bf16":
            if self._shadow is None or not (trust_shadow and self._shadow_fresh):
                if self._shadow is None or self._shadow.device != self._flat.device:
                    self._shadow = torch.empty(self._flat.numel(), dtype=torch.bfloat16, device=self._flat.device)
                L.check(L.lib().gct_cast_f32_to_bf16(L.ptr(self._flat), L.ptr(self._shadow), self._flat.numel(),
                                                    L.stream_ptr()), "gct_cast_f32_to_bf16")
            self._shadow_fresh = bool(trust_shadow)
            shadow = self._shadow
        return L.GctWeights(params_f32=self._flat.data_ptr(), params_bf16=shadow.data_ptr() if shadow is not None else None,
                            grads_f32=grads.data_ptr() if grads is not None else None,
                            slot_offsets_host=self._offsets.ctypes.data)

    # ------------------------------------------------------------------ copy / pickle
    _TRANSIENT = ("_flat", "_shadow", "_ws", "_graph_ws", "_param_list", "_grad_views", "_offsets")

    def __getstate__(self):
        """deepcopy / torch.save(model): the flat buffer, shadow and scratch are rebuilt on the other side."""
        state = self.__dict__.copy()
        for k in self._TRANSIENT:
            state.pop(k, None)
        return state

    def __setstate__(self, state):
        super().__setstate__(state)
        self._shadow, self._shadow_fresh = None, False
        self._ws, self._graph_ws = _Workspace(), []
        self._flatten()

    # ------------------------------------------------------------------ forward / backward
    def _run(self, src, trg, src_mask, trg_mask, econds, dconds, *, run_encoder=True, run_decoder=True, z_in=None,
             eps="auto", want_attn=False):
        lib = L.lib()
        cfg = self._cfg()
        dev = self._flat.device
        nc = int(self.nconds)
        lat = cfg.latent_dim
        if run_encoder:
            L.require_cuda(src, "src")
            src = src.contiguous()
            B, S = src.shape
        else:
            L.require_cuda(z_in, "z")
            z_in = z_in.float().contiguous()
            B, S = z_in.size(0), z_in.size(1) - nc
        Se = nc + S
        T = trg.size(1) if run_decoder else 1
        Ld = T + (nc if (cfg.use_cond2dec and nc > 0) else 0)
        Sm = Se + (nc if (cfg.use_cond2lat and nc > 0 and not cfg.use_cond2dec) else 0)
        if cfg.use_cond2dec and cfg.use_cond2lat and nc > 0:
            raise L.GctError("use_cond2dec together with use_cond2lat is inconsistent in the reference as well "
                             "(mask/memory length mismatch, Model/cvaetf.py:103-116)")
        sm8 = mask_bytes(src_mask, (B, 1, Se)).view(B, Se)
        tm8 = mask_bytes(trg_mask, (B, Ld, Ld)) if run_decoder else None
        train = bool(self.training and torch.is_grad_enabled())
        if run_encoder and isinstance(eps, str):
            eps = torch.randn((B, Se, lat), device=dev, dtype=torch.float32) if self._variational() else None
        f32 = dict(device=dev, dtype=torch.float32)
        logits = torch.empty((B, Ld, cfg.trg_vocab), **f32) if run_decoder else None
        mu = torch.empty((B, Se, lat), **f32) if run_encoder else None
        lv = torch.empty((B, Se, lat), **f32) if run_encoder else None
        z = torch.empty((B, Se, lat), **f32) if run_encoder else None
        H, N = cfg.heads, cfg.n_layers
        attn = [None, None, None]
        if want_attn:
            if run_encoder:
                attn[0] = torch.empty((N, B, H, Se, Se), **f32)
            if run_decoder:
                attn[1] = torch.empty((N, B, H, Ld, Ld), **f32)
                attn[2] = torch.empty((N, B, H, Ld, Sm), **f32)
        if self._step_seed is None:       # dropout stream follows torch.manual_seed (parity is statistical only)
            self._step_seed = torch.initial_seed() & 0xFFFFFFFF
        self._step_seed = (self._step_seed * 1664525 + 1013904223) & 0xFFFFFFFF
        io = dict(src=src if run_encoder else None, trg=trg.contiguous() if run_decoder else None, sm8=sm8, tm8=tm8,
                  econds=_f32c(econds) if (nc > 0 and run_encoder) else None,
                  dconds=_f32c(dconds) if (nc > 0 and run_decoder and dconds is not None) else None, eps=eps if run_encoder else None,
                  z_in=z_in, B=B, S=S, T=T, train=int(train), seed=self._step_seed, run_encoder=int(run_encoder),
                  run_decoder=int(run_decoder), logits=logits, mu=mu, lv=lv, z=z, attn=attn)
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self._param_list)
        ws_bytes = lib.gct_forward_workspace_bytes(C.byref(cfg), B, S, T)
        if need_grad:
            ws = torch.empty(int(ws_bytes), dtype=torch.uint8, device=dev)      # owned by the autograd node
            outs = _ModelFn.apply(self, cfg, io, ws, *self._param_list)
            logits, mu, lv, z = outs
        else:
            ws = self._ws.get("fwd", ws_bytes, dev)
            self._forward_raw(cfg, io, ws)
        return logits, mu, lv, z, attn

    def _io_struct(self, io) -> L.GctIO:
        a = io["attn"]
        return L.GctIO(src=_p(io["src"]), trg=_p(io["trg"]), src_mask=_p(io["sm8"]), trg_mask=_p(io["tm8"]),
                       econds=_p(io["econds"]), dconds=_p(io["dconds"]), eps=_p(io["eps"]), z_in=_p(io["z_in"]), B=io["B"],
                       S=io["S"], T=io["T"], train=io["train"], seed=io["seed"], run_encoder=io["run_encoder"],
                       run_decoder=io["run_decoder"], logits=_p(io["logits"]), mu=_p(io["mu"]), log_var=_p(io["lv"]),
                       z=_p(io["z"]), enc_attn=_p(a[0]), dec_attn1=_p(a[1]), dec_attn2=_p(a[2]))

    def _forward_raw(self, cfg, io, ws):
        w = self._weights()
        ios = self._io_struct(io)
        L.check(L.lib().gct_forward(C.byref(cfg), C.byref(w), C.byref(ios), L.ptr(ws), ws.numel(), L.stream_ptr()), "gct_forward")

    def _backward_raw(self, cfg, io, ws, grads_flat, dlogits, dmu, dlv, dz):
        lib = L.lib()
        w = self._weights(grads_flat)
        ios = self._io_struct(io)
        sb = lib.gct_backward_scratch_bytes(C.byref(cfg), io["B"], io["S"], io["T"])
        scratch = self._ws.get("bwd", sb, ws.device)
        L.check(lib.gct_backward(C.byref(cfg), C.byref(w), C.byref(ios), _p(dlogits), _p(dmu), _p(dlv), _p(dz), L.ptr(ws),
                                 ws.numel(), L.ptr(scratch), scratch.numel(), L.stream_ptr()), "gct_backward")

    def _variational(self) -> bool:
        holder = self.sampler if hasattr(self, "sampler") else self.encoder
        return bool(getattr(holder, "variational", True))

    def grad_views(self, grads_flat):
        out = []
        for off, n, shape in self._grad_views:
            out.append(grads_flat[off:off + n].view(shape))
        return out


def _p(t):
    return None if t is None else t.data_ptr()


def _f32c(t):
    if t is None:
        return None
    L.require_cuda(t, "conditions")
    return t.float().contiguous()


def mask_bytes(mask, shape):
    from .modules import mask_to_bytes
    if mask is None:
        raise L.GctError("mask missing")
    return mask_to_bytes(mask, shape)


class _ModelFn(torch.autograd.Function):
    """Whole-network autograd node: forward = gct_forward, backward = gct_backward."""

    @staticmethod
    def forward(ctx, model, cfg, io, ws, *params):
        model._forward_raw(cfg, io, ws)
        ctx.model, ctx.cfg, ctx.io, ctx.ws = model, cfg, io, ws
        outs = tuple(t for t in (io["logits"], io["mu"], io["lv"], io["z"]))
        ctx.present = [t is not None for t in outs]
        ctx.set_materialize_grads(False)
        return outs

    @staticmethod
    def backward(ctx, dlogits, dmu, dlv, dz):
        model = ctx.model

        def prep(g):
            return None if g is None else g.float().contiguous()

        grads_flat = torch.zeros_like(model._flat)
        model._backward_raw(ctx.cfg, ctx.io, ctx.ws, grads_flat, prep(dlogits), prep(dmu), prep(dlv), prep(dz))
        views = model.grad_views(grads_flat)
        ctx.ws = None
        return (None, None, None, None) + tuple(v if p.requires_grad else None for v, p in zip(views, model._param_list))
