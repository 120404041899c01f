"""Drop-in for Inference/sampling_tool.py: same classes, kwargs and return values, but
``Sampling.decode`` runs a KV-cached decoder on the device (gct_decode_begin / gct_decode_steps)
instead of re-running the whole decoder on the growing prefix every step.

What stays identical to the reference (sampling_tool.py:140-184): one token per step for every
row, finished rows are not frozen, the loop stops after the first step at which every row has
emitted <eos> at least once (checked every ``sync_every`` steps; the tail is cut off afterwards so
the returned ``ys`` has exactly the reference's length), greedy = first maximal index.
Multinomial draws come from the device (inverse CDF on U(0,1)) -- statistically, not bitwise,
equal to torch.multinomial.
"""
from __future__ import annotations

import ctypes as C
import time
from collections import OrderedDict

import numpy as np
import torch

from .. import _lib as L
from ..Model.modules import get_src_mask, get_trg_mask
from .toklen_sampling import tokenlen_gen_from_data_distribution


_MT_OK = None


def _mt_layout_ok() -> bool:
    """One-off check that torch.get_rng_state() still has the layout gct_mt19937_fill assumes (seed u64, left i32, seeded i32,
    next u64, 624 x u64 words, ... = 5056 bytes) AND that the engine semantics match: 40 uniforms produced from a copy of the
    state must equal torch.rand's.  The global generator is left untouched."""
    global _MT_OK
    if _MT_OK is None:
        try:
            keep = torch.get_rng_state()
            ok = keep.numel() == 5056
            if ok:
                sn = keep.clone().numpy()
                raw = np.empty(700, dtype=np.uint32)
                L.check(L.lib().gct_mt19937_fill(sn[24:24 + 4992].view(np.uint64).ctypes.data, sn[8:12].view(np.int32).ctypes.data,
                                                 sn[16:24].view(np.uint64).ctypes.data, raw.ctypes.data, 700), "gct_mt19937_fill")
                mine = (raw & 0xFFFFFF).astype(np.float32) * np.float32(2.0 ** -24)
                theirs = torch.rand(700).numpy()
                after = torch.get_rng_state().numpy()
                ok = bool(np.array_equal(mine, theirs)) and bool(np.array_equal(after[:5016], sn[:5016]))
            torch.set_rng_state(keep)
            _MT_OK = ok
        except Exception:
            _MT_OK = False
    return _MT_OK


class Sampling:
    def __init__(self, model, kwargs, top_k=None):
        self.batch_size = 512
        self.model = model
        self.top_k = top_k          # dead in the reference too: no subclass passes it

        self.SRC = kwargs['SRC']
        self.TRG = kwargs['TRG']
        self.pad_id = self.SRC.vocab.stoi['<pad>']
        self.sos_id = self.TRG.vocab.stoi['<sos>']
        self.eos_id = self.TRG.vocab.stoi['<eos>']
        self.sep_id = self.TRG.vocab.stoi['<sep>'] if '<sep>' in self.TRG.vocab.stoi else None

        self.cond_dim = kwargs['cond_dim']
        self.latent_dim = kwargs['latent_dim']
        self.max_strlen = kwargs['max_strlen']
        self.use_cond2dec = kwargs['use_cond2dec']
        self.decode_algo = kwargs['decode_algo']
        self.toklen_data = kwargs['toklen_data']
        self.scaler = kwargs['scaler']
        self.device = kwargs['device']
        self.n_jobs = kwargs['n_jobs']

        # knobs that do not exist in the reference
        self.sync_every = kwargs.get('sync_every', 16)        # steps between <eos> checks (one D2H read each)
        self.use_cuda_graph = kwargs.get('use_cuda_graph', True)
        self.latent_bucket = kwargs.get('latent_bucket', 8)   # latent length padded (masked) to a multiple of this
        # sample_smiles with >= pipeline_rows rows runs as two halves: the second half's host->device copy and the first half's
        # detokenisation overlap the other half's decode.  None = decide from the measured host cost per row: worth it with a
        # regex tokeniser in TRG.tokenize (~10 us / row), not with a trivial one ([B200], 30k rows: 40.2k vs 41.6k SMILES/s)
        self.pipeline_rows = kwargs.get('pipeline_rows', None)
        self.z_on_device = kwargs.get('z_on_device', False)
        self.host_z_exact = kwargs.get('host_z_exact', False)
        self.decode_streams = kwargs.get('decode_streams', None)   # row groups decoded on concurrent streams (None = by batch size)
        # Active-row decode (opt-in; the reference re-runs EVERY row until the last one has emitted <eos>, :144-183): rows that
        # have emitted <eos> append <pad> from then on, their attention work is skipped, and -- for calls of at least
        # compact_min_rows rows -- every compact_every steps the rows still running are gathered through a row map so that the
        # GEMMs / Norms shrink with them (batch rounded up to compact_quantum rows).  The strings sample_smiles returns are
        # unchanged (id_to_smi stops at the first <eos>); decode()'s ys differs from the reference's AFTER each row's <eos>.
        self.skip_finished = kwargs.get('skip_finished', False)
        self.compact_every = kwargs.get('compact_every', 8)
        self.compact_quantum = kwargs.get('compact_quantum', None)      # None: n / 32 rounded up to a multiple of 256
        self.compact_min_rows = kwargs.get('compact_min_rows', 4096)
        self.last_row_steps = 0          # sum over the steps run of the rows the step kernels worked on
        self.last_steps_executed = 0     # steps the device ran (whole chunks), >= last_decode_steps
        self._group_streams = []
        self._host_s_per_row = 0.0
        self._side_stream = None
        # static decode buffers / captured CUDA graphs per request shape: small LRUs (each entry pins O(n * Lz * latent)
        # floats of HBM and a graph per chunk), graphs of a superseded workspace are dropped
        self.cache_shapes = kwargs.get('cache_shapes', 4)
        self._graphs = OrderedDict()
        self._static = OrderedDict()
        self._itos = np.array(self.TRG.vocab.itos, dtype=object)
        model.pad_id = self.pad_id
        self.last_decode_steps = 0

    # ------------------------------------------------------------------ small helpers (reference :43-137)
    def init_y(self, n, add_sos=True, sca_ids=None, add_sep=False):
        start_ids = []
        if add_sos:
            start_ids.append(self.sos_id)
        if sca_ids is not None:
            start_ids.extend(sca_ids)
        if add_sep:
            start_ids.append(self.sep_id)
        return torch.from_numpy(np.stack([start_ids] * n))

    def id_to_smi(self, ids):
        smi = ''
        for i in ids:
            if i == self.eos_id:
                break
            if i != self.sos_id:
                smi += self.TRG.vocab.itos[i]
        return smi

    def ids_to_smiles(self, outs: np.ndarray):
        """Batch version of id_to_smi (Inference/sampling_tool.py:54-61 of the reference): cut each row at its first
        <eos>, drop <sos>, join the token strings.  One pass in the library's host-side detokeniser (gct_detokenize)
        writes newline-terminated rows; Python only decodes and splits the blob."""
        outs = np.ascontiguousarray(outs, dtype=np.int16)
        n, width = outs.shape
        vt = self._vocab_table()
        if vt is None or n == 0:              # a token contains a newline: row-by-row join
            itos, sos, eos = self._itos, self.sos_id, self.eos_id
            is_eos = outs == eos
            end = np.where(is_eos.any(axis=1), is_eos.argmax(axis=1), width)
            return [''.join(itos[row[:e][row[:e] != sos]]) for row, e in zip(outs, end)]
        blob, voff, maxlen = vt
        out = np.empty(n * (width * maxlen + 1), dtype=np.uint8)
        nb = L.lib().gct_detokenize(outs.ctypes.data, n, width, blob.ctypes.data, voff.ctypes.data, len(voff) - 1, int(self.eos_id),
                                    int(self.sos_id), out.ctypes.data, out.size)
        if nb < 0:
            L.check(int(nb), "gct_detokenize")
        return out[:nb].tobytes().decode('utf-8').split('\n')[:-1]

    def _vocab_table(self):
        vt = getattr(self, '_vocab_tab', False)
        if vt is False:
            enc = [str(t).encode('utf-8') for t in self._itos]
            if any(b'\n' in e for e in enc):
                vt = None
            else:
                voff = np.zeros(len(enc) + 1, dtype=np.int32)
                voff[1:] = np.cumsum([len(e) for e in enc])
                blob = np.frombuffer(b''.join(enc) + b'\x00', dtype=np.uint8).copy()
                vt = (blob, voff, max(1, max(len(e) for e in enc)))
            self._vocab_tab = vt
        return vt

    def smi_to_id(self, smi, add_sos=False, add_sep=False, add_eos=False):
        ids = []
        if add_sos:
            ids.append(self.sos_id)
        if add_sep:
            ids.append(self.sep_id)
        ids.extend(self.TRG.vocab.stoi[t] for t in self.TRG.tokenize(smi))
        if add_eos:
            ids.append(self.eos_id)
        return ids

    def sample_toklen(self, n):
        n_bin = int(self.toklen_data.max() - self.toklen_data.min())
        toklens = tokenlen_gen_from_data_distribution(data=self.toklen_data, size=n, nBins=n_bin)
        toklens = toklens.reshape((-1,)) + self.cond_dim
        return np.rint(toklens).astype(int)

    def tokenize_smiles(self, smiles_list, field='SRC'):
        if field != 'SRC':
            raise ValueError('only the SRC field is tokenised by the reference')
        p = self.SRC.process([self.SRC.tokenize(smi) for smi in smiles_list])
        return p if self.SRC.batch_first else p.T

    def sample_z(self, toklen, n):
        """z ~ N(0, 1) from the HOST generator like the reference (Inference/sampling_tool.py:93-97), so that a seeded run
        draws the same latents.  torch.normal on the CPU costs ~7 ns per element (2 s for 30 000 x 55 x 128 -- three times
        the whole decode), so the draw is split: the library advances torch's CPU MT19937 engine in a tight host loop
        (gct_mt19937_fill, one raw output per element into pinned memory, ~1.5 ns each) and the device applies
        at::normal_fill's Box-Muller blocks (gct_normal_from_mt).  Same generator state afterwards, same values to the
        rounding of logf / sincosf (1-2 ulp); returns a CUDA tensor.  `host_z_exact=True` (sampler kwarg) keeps
        torch.normal itself; `z_on_device=True` draws with the CUDA generator instead (not seed-compatible)."""
        if self.z_on_device:
            return torch.randn((n, toklen, self.latent_dim), device=self.device)
        numel = int(n) * int(toklen) * int(self.latent_dim)
        if self.host_z_exact or numel < 16 or not torch.device(self.device).type == 'cuda' or not _mt_layout_ok():
            return torch.normal(mean=0, std=1, size=(n, toklen, self.latent_dim))
        lib = L.lib()
        extra = 16 if numel % 16 else 0
        raw = getattr(self, '_raw_host', None)
        if raw is None or raw.numel() < numel + extra:          # grow-only pinned staging buffer, 25 % headroom: latent lengths vary call to call
            raw = self._raw_host = torch.empty(int((numel + extra) * 1.25), dtype=torch.int32, pin_memory=True)
        state = torch.get_rng_state()
        sn = state.numpy()
        words = sn[24:24 + 624 * 8].view(np.uint64)
        left, nxt = sn[8:12].view(np.int32), sn[16:24].view(np.uint64)
        L.check(lib.gct_mt19937_fill(words.ctypes.data, left.ctypes.data, nxt.ctypes.data, raw.data_ptr(), numel + extra),
                "gct_mt19937_fill")
        torch.set_rng_state(state)
        raw_dev = raw[:numel + extra].to(self.device, non_blocking=True)
        z = torch.empty((n, toklen, self.latent_dim), device=self.device, dtype=torch.float32)
        L.check(lib.gct_normal_from_mt(L.ptr(raw_dev), L.ptr(z), numel, L.stream_ptr()), "gct_normal_from_mt")
        return z

    def transform(self, prop):
        return torch.from_numpy(self.scaler.transform(prop)).float()

    def encoder_input(self, smiles_list, transform=False, econds=None):
        kwargs = {'src': self.tokenize_smiles(smiles_list).to(self.device)}
        if econds is not None:
            if transform:
                econds = self.transform(econds)
            if not torch.is_tensor(econds):
                econds = torch.from_numpy(np.array(econds))
            kwargs['econds'] = econds.float().to(self.device)
        return kwargs

    def decoder_input(self, ys, z, dconds=None, transform=False):
        kwargs = {'z': z.to(self.device), 'trg': ys.to(self.device)}
        if dconds is not None:
            if transform:
                dconds = self.transform(dconds)
            if not torch.is_tensor(dconds):
                dconds = torch.from_numpy(np.array(dconds))
            kwargs['dconds'] = dconds.to(self.device)
        return kwargs

    # ------------------------------------------------------------------ the decode loop
    def decode(self, **kwargs):
        """kwargs: zs (n,Lz,lat), ys (n,t0) int64 prefix, src_mask (n,1,Lz) bool[, dconds (n,nc)].
        Returns ys (n, t0 + steps_run) exactly like the reference's loop."""
        if self.use_cond2dec and self.cond_dim > 0:
            return self._decode_recompute(**kwargs)
        with torch.no_grad():
            return self._decode_cached(**kwargs)

    def _decode_recompute(self, **kwargs):
        """use_cond2dec puts non-causal condition rows in front of the target, which a per-token KV
        cache cannot express; this path re-runs model.decode on the prefix like the reference."""
        ys = kwargs['ys'].to(self.device)
        done = torch.zeros(ys.size(0), dtype=torch.bool)
        with torch.no_grad():
            for _ in range(self.max_strlen - 1):
                trg_mask = get_trg_mask(ys, self.pad_id, self.use_cond2dec, kwargs['dconds'])
                out = self.model.decode(trg=ys, z=kwargs['zs'].to(self.device), src_mask=kwargs['src_mask'],
                                        trg_mask=trg_mask, dconds=kwargs['dconds'])
                prob = torch.softmax(out[:, self.cond_dim:, :][:, -1, :].float(), dim=-1)
                if self.decode_algo == 'greedy':
                    nxt = prob.max(dim=1)[1]
                else:
                    nxt = torch.multinomial(prob, 1).squeeze(-1)
                ys = torch.cat([ys, nxt.unsqueeze(-1)], dim=1)
                done |= (nxt.cpu() == self.eos_id)
                if bool(done.all()):
                    break
        return ys

    def _row_groups(self, n, probe):
        """How many independent row groups a decode of n rows runs as (`decode_streams` sampler kwarg, default 1).  Rows are
        independent, so a batch can be cut into groups whose kernel chains run on concurrent streams (one CUDA graph with
        parallel branches per chunk of steps).  [B200] it does not pay at any batch size (profiles/r02_ab_streams.txt: batch 512
        46.2 ms as one group, 54.6 as two, 55.3 as four; batch 128: 38.9 / 41.6 / 52.5): a step costs ~37 ms / 99 of fixed
        per-kernel latency whatever the rows, but the small-tile GEMMs already occupy every SM's CTA slots (shared memory),
        so the branches serialise instead of overlapping.  Kept as an option for other shapes; the default is one group."""
        if probe or not self.decode_streams:
            return 1
        return max(1, min(int(self.decode_streams), n))

    def _decode_cached(self, zs, ys, src_mask, dconds=None, uniforms=None, idle_work=None, forced=None, probs_out=None,
                       logits_out=None):
        lib, model, dev = L.lib(), self.model, torch.device(self.device)
        cfg = model._cfg()
        n, t0 = ys.shape
        steps = self.max_strlen - 1
        max_len = t0 + steps
        Lz = zs.size(1)
        Lzp = (Lz + self.latent_bucket - 1) // self.latent_bucket * self.latent_bucket
        greedy = int(self.decode_algo == 'greedy')
        nc = self.cond_dim
        probe = forced is not None or probs_out is not None or logits_out is not None
        G = self._row_groups(n, probe)
        bounds = [(n * g // G, n * (g + 1) // G) for g in range(G)]

        skip = bool(self.skip_finished) and not probe
        compacting = skip and G == 1 and n >= self.compact_min_rows
        key = (n, Lzp, max_len, t0, greedy, nc, model.compute_dtype, G, skip)
        st = self._static.get(key)
        if st is None:
            st = dict(zs=torch.zeros((n, Lzp, self.latent_dim), device=dev, dtype=torch.float32),
                      mask=torch.zeros((n, Lzp), device=dev, dtype=torch.uint8),
                      ys=torch.zeros((n, max_len), device=dev, dtype=torch.int64),
                      status=torch.zeros((G, 2), device=dev, dtype=torch.int32),
                      uni=[torch.zeros((steps, hi - lo), device=dev, dtype=torch.float32) for lo, hi in bounds],
                      dconds=torch.zeros((n, max(nc, 1)), device=dev, dtype=torch.float32),
                      status_host=torch.zeros((G, 2), dtype=torch.int32).pin_memory())
            self._static[key] = st
            while len(self._static) > self.cache_shapes:
                old, _ = self._static.popitem(last=False)
                for gk in [gk for gk in self._graphs if gk[:len(old)] == old]:
                    del self._graphs[gk]
        else:
            self._static.move_to_end(key)
        st['zs'][:, :Lz].copy_(zs, non_blocking=True)
        if Lzp > Lz:
            st['zs'][:, Lz:].zero_()
            st['mask'][:, Lz:].zero_()
        st['mask'][:, :Lz].copy_(src_mask.reshape(n, Lz).to(torch.uint8), non_blocking=True)
        st['ys'][:, :t0].copy_(ys, non_blocking=True)
        if skip:
            st['ys'][:, t0:].fill_(int(self.pad_id))      # rows gathered away after their <eos> are not written any more
        if nc > 0:
            st['dconds'].copy_(dconds.float(), non_blocking=True)
        if not greedy:
            for (lo, hi), u in zip(bounds, st['uni']):
                if uniforms is not None:
                    u.copy_(uniforms[:, lo:hi], non_blocking=True)
                else:
                    u.uniform_()
        w = model._weights()
        if forced is not None:
            assert forced.is_cuda and forced.dtype == torch.int64 and tuple(forced.shape) == (n, max_len) and forced.is_contiguous()
        for t in (probs_out, logits_out):
            assert t is None or (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
                                 and tuple(t.shape) == (steps, n, cfg.trg_vocab))
        wss, decs = [], []
        for g, (lo, hi) in enumerate(bounds):
            ng = hi - lo
            ws = model._ws.get(('decode', g), lib.gct_decode_workspace_bytes(C.byref(cfg), ng, Lzp, max_len), dev)
            wss.append(ws)
            decs.append(L.GctDecode(B=ng, Lz=Lzp, max_len=max_len, prefix_len=t0, greedy=greedy, eos_id=int(self.eos_id), seed=0,
                                    zs=st['zs'][lo:].data_ptr(), src_mask=st['mask'][lo:].data_ptr(),
                                    dconds=st['dconds'][lo:].data_ptr() if nc > 0 else None,
                                    uniforms=None if greedy else st['uni'][g].data_ptr(), ys=st['ys'][lo:].data_ptr(),
                                    status=st['status'][g].data_ptr(), forced=L._p(forced), probs_out=L._p(probs_out),
                                    logits_out=L._p(logits_out), skip_done=int(skip), n_active=0, rowmap=None))
        if G > 1 and len(self._group_streams) < G - 1:
            self._group_streams += [torch.cuda.Stream(device=dev) for _ in range(G - 1 - len(self._group_streams))]

        def on_groups(fn):
            """fn(g) for every row group: group 0 on the current stream, the others forked onto side streams and joined back."""
            cur = torch.cuda.current_stream(dev)
            for g in range(1, G):                       # fork BEFORE any group's work is enqueued on the current stream
                self._group_streams[g - 1].wait_stream(cur)
            fn(0)
            for g in range(1, G):
                with torch.cuda.stream(self._group_streams[g - 1]):
                    fn(g)
            for g in range(1, G):
                cur.wait_stream(self._group_streams[g - 1])

        def begin():
            on_groups(lambda g: L.check(lib.gct_decode_begin(C.byref(cfg), C.byref(w), C.byref(decs[g]), L.ptr(wss[g]), wss[g].numel(),
                                                             L.stream_ptr()), "gct_decode_begin"))

        def run(s0, s1):
            on_groups(lambda g: L.check(lib.gct_decode_steps(C.byref(cfg), C.byref(w), C.byref(decs[g]), s0, s1, L.ptr(wss[g]),
                                                             wss[g].numel(), L.stream_ptr()), "gct_decode_steps"))

        chunks = [(s, min(steps, s + self.sync_every)) for s in range(0, steps, self.sync_every)]
        wptrs = tuple(ws.data_ptr() for ws in wss)
        gkey = key + (wptrs, w.params_f32, w.params_bf16, self.sync_every)
        for gk in [gk for gk in self._graphs if gk[:len(key)] == key and gk[len(key)] != wptrs]:
            del self._graphs[gk]                # a grow-only workspace was reallocated: those graphs point at freed memory
        # probe buffers are per call: not baked into a graph; a compacting decode changes its batch between chunks and has
        # >= compact_min_rows rows per step kernel, so the host stays ahead of the device without graphs
        use_graph = self.use_cuda_graph and not probe and not compacting
        graphs = self._graphs.get(gkey) if use_graph else None
        if use_graph and graphs is None and st.get('warm'):
            # second call with this shape: capture begin + every chunk once, replay from now on
            graphs = []
            for fn in [begin] + [(lambda a=a, b=b: run(a, b)) for a, b in chunks]:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    fn()
                graphs.append(g)
            self._graphs[gkey] = graphs
            # capture does not execute: fall through to replay below
        st['warm'] = True
        steps_run = steps

        def all_done():
            st['status_host'].copy_(st['status'], non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return int(st['status_host'][:, 0].sum()) >= n      # every group counts its own rows that have emitted <eos>

        if graphs is not None:
            graphs[0].replay()
        else:
            begin()
        row_steps = executed = 0
        if compacting:
            q = self.compact_quantum or max(256, -(-n // (32 * 256)) * 256)
            rowmap = st.get('rowmap')
            if rowmap is None:
                rowmap = st['rowmap'] = torch.empty(n, device=dev, dtype=torch.int32)
            dec, ws, rows = decs[0], wss[0], n
            chunks = []                                   # the loop below replaces the chunked one
            s0 = 0
            while s0 < steps:
                s1 = min(steps, s0 + max(1, int(self.compact_every)))
                run(s0, s1)
                row_steps += rows * (s1 - s0)
                executed += s1 - s0
                if idle_work is not None:
                    idle_work.step()
                s0 = s1
                if s0 < steps:
                    if all_done():
                        break
                    n_act = n - int(st['status_host'][0, 0])
                    want = min(n, max(q, -(-n_act // q) * q))
                    if want < rows:
                        L.check(lib.gct_decode_compact(C.byref(cfg), C.byref(dec), want, L.ptr(rowmap), L.ptr(ws), ws.numel(),
                                                       L.stream_ptr()), "gct_decode_compact")
                        dec.n_active, dec.rowmap, rows = want, rowmap.data_ptr(), want
        for ci, (s0, s1) in enumerate(chunks):
            if graphs is not None:
                graphs[1 + ci].replay()
            else:
                run(s0, s1)
            row_steps += n * (s1 - s0)
            executed += s1 - s0
            if idle_work is not None:
                idle_work.step()          # bounded slice of host work while the GPU runs this chunk
            if s1 < steps and all_done():
                break
        if all_done():
            steps_run = int(st['status_host'][:, 1].max()) + 1      # the step at which the LAST group completed
        self.last_decode_steps = steps_run
        self.last_steps_executed = executed
        self.last_row_steps = row_steps           # rows x steps the step kernels ran on (chunks run to their end: >= n * steps_run
                                                  # for a plain decode that stopped early inside a chunk)
        return st['ys'][:, :t0 + steps_run].clone()

    def teacher_forced_logits(self, zs, ys_full, src_mask, dconds=None, t0=1, want_probs=False):
        """Runs the KV-cached decoder over GIVEN sequences: ys_full (n, t0 + max_strlen - 1) holds the prefix (t0 tokens)
        and the tokens to append at every step.  Returns logits (steps, n, Vt) -- row s equals
        model.decode(ys_full[:, :t0+s], ...)[:, -1] of the reference loop (Inference/sampling_tool.py:150-160) --
        and optionally their softmax.  Not in the reference; used to score sequences and by the per-step parity tests."""
        dev = torch.device(self.device)
        n, steps = ys_full.size(0), self.max_strlen - 1
        assert ys_full.size(1) == t0 + steps
        V = self.model.out.out_features
        forced = ys_full.to(dev).long().contiguous()
        logits = torch.empty((steps, n, V), device=dev, dtype=torch.float32)
        probs = torch.empty_like(logits) if want_probs else None
        with torch.no_grad():
            self._decode_cached(zs.to(dev), forced[:, :t0].contiguous(), src_mask.to(dev), dconds=dconds, forced=forced,
                                probs_out=probs, logits_out=logits)
        return (logits, probs) if want_probs else logits

    # ------------------------------------------------------------------ shared tail of sample_smiles
    def _finish(self, outs, strip):
        outs = outs.to(torch.int16).cpu().numpy()          # ids < 2^15: a quarter of the int64 device->host bytes
        t0 = time.perf_counter()
        smiles = self.ids_to_smiles(outs[:, strip:])
        toklen_gen = list(map(len, map(self.TRG.tokenize, smiles)))
        self._host_s_per_row = (time.perf_counter() - t0) / max(1, len(smiles))
        return smiles, toklen_gen

    class _Detok:
        """Detokenises a decoded half in slices, called from the decode loop of the other half."""

        def __init__(self, owner, outs_dev, strip, rows_per=4096):
            self.o, self.strip, self.rows_per = owner, strip, rows_per
            self.host = torch.empty(outs_dev.shape, dtype=torch.int16).pin_memory()
            self.host.copy_(outs_dev.to(torch.int16), non_blocking=True)
            self.event = torch.cuda.Event()
            self.event.record()
            self.i, self.smiles, self.toklen = 0, [], []

        def step(self):
            if self.i >= self.host.size(0) or not self.event.query():
                return
            sl = self.host.numpy()[self.i:self.i + self.rows_per, self.strip:]
            s = self.o.ids_to_smiles(sl)
            self.smiles += s
            self.toklen += list(map(len, map(self.o.TRG.tokenize, s)))
            self.i += self.rows_per

        def finish(self):
            self.event.synchronize()
            while self.i < self.host.size(0):
                self.step()
            return self.smiles, self.toklen

    def _decode_finish(self, strip, zs, ys, src_mask, dconds=None):
        """decode + detokenise for sample_smiles.  Large requests (>= pipeline_rows) run as two halves so that the host side
        of one half (pinned host->device copy of its latents, detokenisation of its result) overlaps the decode of the other;
        rows are independent, so the result equals the single-pass one row for row (up to the multinomial draw assignment)."""
        dev, n = self.device, ys.size(0)
        if self.pipeline_rows is not None:
            pipelined = n >= self.pipeline_rows
        else:
            pipelined = n >= 8192 and self._host_s_per_row * n >= 0.1        # >= 100 ms of host work to hide
        if not pipelined or (self.use_cond2dec and self.cond_dim > 0):
            kw = dict(zs=zs.to(dev, non_blocking=True), ys=ys.to(dev), src_mask=src_mask.to(dev))
            if dconds is not None:
                kw['dconds'] = dconds.to(dev)
            return self._finish(self.decode(**kw), strip)
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream(device=dev)
        cur, side, h = torch.cuda.current_stream(dev), self._side_stream, (n + 1) // 2

        def stage(lo, hi, stream):
            with torch.cuda.stream(stream):
                kw = dict(zs=zs[lo:hi].to(dev, non_blocking=True), ys=ys[lo:hi].to(dev, non_blocking=True),
                          src_mask=src_mask[lo:hi].to(dev, non_blocking=True))
                if dconds is not None:
                    kw['dconds'] = dconds[lo:hi].to(dev, non_blocking=True)
            return kw

        kw_a = stage(0, h, cur)
        kw_b = stage(h, n, side)                       # copies run beside the first half's decode
        outs_a = self.decode(**kw_a)
        steps_a = self.last_decode_steps
        work = Sampling._Detok(self, outs_a, strip)
        cur.wait_stream(side)
        for t in kw_b.values():
            t.record_stream(cur)
        with torch.no_grad():
            outs_b = self._decode_cached(**kw_b, idle_work=work)
        self.last_decode_steps = max(steps_a, self.last_decode_steps)
        smiles_a, tok_a = work.finish()
        smiles_b, tok_b = self._finish(outs_b, strip)
        return smiles_a + smiles_b, tok_a + tok_b

    def _latent_mask(self, toklen, n, width, offset=0):
        stop = torch.LongTensor(np.asarray(toklen)).view(n, 1, 1) + offset
        return torch.arange(width).expand(n, 1, width) < stop

    def _attention_maps(self, src, src_mask, ys, econds=None, dconds=None):
        """encoder self-attention, decoder self- and cross-attention probabilities (lists of N
        tensors (n,H,Lq,Lk)) with z = mu, as get_attention_map of the reference computes them."""
        m = self.model
        _, mu, _, _, att = m._run(src, None, src_mask, None, econds, None, run_decoder=False, want_attn=True)
        n, Lz = mu.size(0), mu.size(1)
        smask = torch.ones((n, 1, Lz), dtype=torch.bool, device=mu.device)
        trg_mask = get_trg_mask(ys, self.pad_id, self.use_cond2dec, dconds)
        _, _, _, _, att2 = m._run(None, ys, smask, trg_mask, None, dconds, run_encoder=False, z_in=mu, want_attn=True)
        return list(att[0]), list(att2[1]), list(att2[2])


class VaetfSampling(Sampling):
    def get_attention_map(self, smiles):
        kws = self.encoder_input([smiles])
        ids = [self.TRG.vocab.stoi[t] for t in ['<sos>'] + self.TRG.tokenize(smiles) + ['<eos>']]
        ys = torch.tensor([ids], dtype=torch.long, device=self.device)
        return self._attention_maps(kws['src'], get_src_mask(kws['src'], self.pad_id), ys)

    def encode_smiles(self, smiles_list):
        kwargs = self.encoder_input(smiles_list)
        kwargs['src_mask'] = get_src_mask(kwargs['src'], self.pad_id)
        return self.model.encode(**kwargs)

    def encode_batch(self, batch):
        batch['src'] = batch['src'].to(self.device)
        return self.model.encode(src=batch['src'], src_mask=get_src_mask(batch['src'], self.pad_id))

    def sample_smiles(self, n, zs=None, toklen=None):
        ys = self.init_y(n, add_sos=True)
        if zs is not None:
            assert n == zs.size(0)
            if toklen is None:
                toklen = [zs.size(1)] * zs.size(0)
        elif toklen is None:
            toklen = self.sample_toklen(n)
        max_toklen = max(toklen)
        if zs is None:
            zs = self.sample_z(max_toklen, n)
        src_mask = self._latent_mask(toklen, n, max_toklen)
        smiles, toklen_gen = self._decode_finish(0, zs, ys, src_mask)
        return smiles, toklen, toklen_gen


class CvaetfSampling(Sampling):
    def encode_smiles(self, smiles_list, econds, transform=True):
        kwargs = self.encoder_input(smiles_list, transform, econds)
        kwargs['src_mask'] = get_src_mask(kwargs['src'], self.pad_id, kwargs['econds'])
        return self.model.encode(**kwargs)

    def encode_batch(self, batch, transform=True):
        if transform:
            batch['econds'] = self.transform(batch['econds'].cpu())
        batch['src'] = batch['src'].to(self.device)
        batch['econds'] = batch['econds'].to(self.device)
        batch['src_mask'] = get_src_mask(batch['src'], self.pad_id, batch['econds'])
        return self.model.encode(**batch)

    def sample_smiles(self, dconds, zs=None, toklen=None, transform=True):
        if zs is not None:
            assert len(dconds) == len(zs), "The number of 'dconds' and 'zs' should be the same!"
        n = len(dconds)
        ys = self.init_y(n, add_sos=True)
        if transform:
            dconds = self.transform(dconds)
        if zs is not None:
            if toklen is None:
                toklen = [zs.size(1)] * zs.size(0)
        elif toklen is None:
            toklen = self.sample_toklen(n)
        else:
            toklen = [t + self.cond_dim for t in toklen]
        max_toklen = max(toklen)
        if zs is None:
            zs = self.sample_z(max_toklen, n)
        src_mask = self._latent_mask(toklen, n, max_toklen)
        smiles, toklen_gen = self._decode_finish(0, zs, ys, src_mask, torch.as_tensor(dconds).float())
        return smiles, toklen, toklen_gen


class _ScaffoldMixin:
    def _sample_with_scaffold(self, n, scaffold, zs, toklen, dconds=None):
        sca_ids = [self.TRG.vocab.stoi[e] for e in self.TRG.tokenize(scaffold)]
        ys = self.init_y(n, add_sos=True, sca_ids=sca_ids, add_sep=True)
        if zs is not None:
            if toklen is None:
                toklen = [zs.size(1) - len(sca_ids) - 1] * zs.size(0)
        elif toklen is None:
            toklen = self.sample_toklen(n)
        max_toklen = max(toklen)
        lat_toklen = len(sca_ids) + 1 + max_toklen
        if zs is None:
            zs = self.sample_z(lat_toklen, n)
        src_mask = self._latent_mask(toklen, n, lat_toklen, offset=len(sca_ids) + 1)
        smiles, toklen_gen = self._decode_finish(1 + len(sca_ids) + 1, zs, ys, src_mask,
                                                 None if dconds is None else torch.as_tensor(dconds).float())
        return smiles, toklen, toklen_gen


class PscavaetfSampling(_ScaffoldMixin, Sampling):
    def encode_smiles(self, smiles_list, scaffold_list, econds, transform=True):
        concat = [s1 + '<sep>' + s2 for s1, s2 in zip(smiles_list, scaffold_list)]
        kwargs = self.encoder_input(concat, transform, econds)
        kwargs['src_mask'] = get_src_mask(kwargs['src'], self.pad_id, kwargs['econds'])
        return self.model.encode(**kwargs)

    def sample_smiles(self, dconds, scaffold, zs=None, toklen=None, transform=True):
        n = len(dconds)
        if transform:
            dconds = self.transform(dconds)
        return self._sample_with_scaffold(n, scaffold, zs, toklen, dconds)


class ScaVaeSampling(_ScaffoldMixin, Sampling):
    def get_attention_map(self, smiles, scaffold):
        kws = self.encoder_input([smiles + '<sep>' + scaffold])
        toks = ['<sos>'] + self.TRG.tokenize(scaffold + '<sep>' + smiles) + ['<eos>']
        ys = torch.tensor([[self.TRG.vocab.stoi[t] for t in toks]], dtype=torch.long, device=self.device)
        return self._attention_maps(kws['src'], get_src_mask(kws['src'], self.pad_id), ys)

    def encode_smiles(self, smiles_list, scaffold_list, transform=True):
        concat = [s1 + '<sep>' + s2 for s1, s2 in zip(smiles_list, scaffold_list)]
        kwargs = self.encoder_input(concat, transform)
        kwargs['src_mask'] = get_src_mask(kwargs['src'], self.pad_id)
        return self.model.encode(**kwargs)

    def sample_smiles(self, n, scaffold, zs=None, toklen=None):
        return self._sample_with_scaffold(n, scaffold, zs, toklen)


sampling_tool_dict = {
    'vaetf': VaetfSampling,
    'pvaetf': CvaetfSampling,
    'scavaetf': ScaVaeSampling,
    'ctf': CvaetfSampling,
    'pscavaetf': PscavaetfSampling,
}
