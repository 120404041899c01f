"""Input pipeline on the device.

Replaces, for the training loop, the reference's ``SmilesDataset.__getitem__`` + ``collate_fn`` + ``Field.process``
(Utils/dataset.py:251-335, Model/collate_fn.py:4-124, torchtext 0.6 ``Field.pad/numericalize``): there every row is
re-tokenised with a regex on every access, padded in Python and copied host->device per batch (``num_workers=0``).
Here the corpus is tokenised ONCE with the same ``Field.tokenize`` / ``vocab.stoi``, kept in HBM as int16 CSR arrays,
and one CUDA launch (``gct_collate``) assembles a whole batch in the layout ``forward_propagation`` expects:

    {'src': (B,S) int64, 'trg': (B,T) int64, ['econds','dconds': (B,nc) float32]}

    src row :            [scaffold <sep>] smiles <pad>...          (SRC has no init / eos token, Utils/field.py:50)
    trg row : <sos>      [scaffold <sep>] smiles <eos> <pad>...    (TRG: init_token='<sos>', eos_token='<eos>', :51-52)

Row order is produced by torch's own ``RandomSampler`` / ``DistributedSampler`` objects, i.e. exactly what the reference's
``DataLoader`` draws under the same seeds (``DataloaderPreparation.get_dataloader``, Utils/dataset.py:317-335).
``randomize_prob > 0`` (RDKit SMILES randomisation) is outside this path: RDKit is not part of the hot path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
from torch.utils.data import DistributedSampler, RandomSampler, SequentialSampler

from .. import _lib as L

_SCAFFOLD_MODELS = ("scavaetf", "pscavaetf")


def _stoi(field, tok):
    stoi = field.vocab.stoi
    try:
        return int(stoi[tok])
    except KeyError:                      # torchtext's stoi is a defaultdict -> <unk>; plain dicts get the same behaviour
        return int(stoi.get("<unk>", 0))


class TokenisedCorpus:
    """All rows of a dataframe (columns ``src`` [, ``src_scaffold``, ``src_<p>``, ``trg_<p>``]) as CSR id arrays."""

    def __init__(self, dataframe, property_list, SRC, TRG, use_scaffold=False, randomize_prob=0):
        if randomize_prob:
            raise L.GctError("randomize_prob > 0 needs RDKit (Utils/smiles.py randomize_smiles); not on the device path")
        self.SRC, self.TRG = SRC, TRG
        self.property_list = list(property_list)
        self.use_scaffold = bool(use_scaffold)
        n = len(dataframe)
        self.n = n

        def csr(column):
            src_ids, trg_ids, off = [], [], np.zeros(n + 1, dtype=np.int64)
            for i, smi in enumerate(column):
                toks_s, toks_t = SRC.tokenize(smi), TRG.tokenize(smi)
                src_ids.extend(_stoi(SRC, t) for t in toks_s)
                trg_ids.extend(_stoi(TRG, t) for t in toks_t)
                if len(toks_s) != len(toks_t):
                    raise L.GctError("SRC and TRG tokenisers disagree on row %d" % i)
                off[i + 1] = off[i] + len(toks_s)
            return np.asarray(src_ids, dtype=np.int16), np.asarray(trg_ids, dtype=np.int16), off

        self.src_ids, self.trg_ids, self.tok_off = csr(dataframe["src"])
        if self.use_scaffold:
            self.sca_src_ids, self.sca_trg_ids, self.sca_off = csr(dataframe["src_scaffold"])
        else:
            self.sca_src_ids = self.sca_trg_ids = self.sca_off = None
        nc = len(self.property_list)
        self.econds = np.ascontiguousarray(dataframe[[f"src_{p}" for p in self.property_list]].to_numpy(np.float32)) if nc else None
        self.dconds = np.ascontiguousarray(dataframe[[f"trg_{p}" for p in self.property_list]].to_numpy(np.float32)) if nc else None
        self.tok_len = np.diff(self.tok_off)
        self.sca_len = np.diff(self.sca_off) if self.use_scaffold else None
        self._dev = None

    def __len__(self):
        return self.n

    def to(self, device):
        """Uploads the arrays once; returns self."""
        device = torch.device(device)
        if device.type != "cuda":
            raise L.GctError("TokenisedCorpus.to(): the batch assembly runs on a CUDA device (no CPU path)")

        def up(a):
            return None if a is None else torch.from_numpy(a).to(device)

        # empty arrays still need a valid pointer
        pad16 = np.zeros(1, dtype=np.int16)
        self._dev = dict(device=device,
                         src_ids=up(self.src_ids if self.src_ids.size else pad16), trg_ids=up(self.trg_ids if self.trg_ids.size else pad16),
                         tok_off=up(self.tok_off),
                         sca_src_ids=up(None if self.sca_off is None else (self.sca_src_ids if self.sca_src_ids.size else pad16)),
                         sca_trg_ids=up(None if self.sca_off is None else (self.sca_trg_ids if self.sca_trg_ids.size else pad16)),
                         sca_off=up(self.sca_off), econds=up(self.econds), dconds=up(self.dconds))
        return self

    def _struct(self):
        d = self._dev
        if d is None:
            raise L.GctError("TokenisedCorpus: call .to(device) first")
        p = lambda t: None if t is None else t.data_ptr()   # noqa: E731
        return L.GctCorpus(src_ids=p(d["src_ids"]), trg_ids=p(d["trg_ids"]), tok_off=p(d["tok_off"]), sca_src_ids=p(d["sca_src_ids"]),
                           sca_trg_ids=p(d["sca_trg_ids"]), sca_off=p(d["sca_off"]), econds=p(d["econds"]), dconds=p(d["dconds"]),
                           nconds=len(self.property_list), n_rows=self.n)

    def batch_shape(self, rows, scaffold_prefix):
        """(S, T) of the padded batch: torchtext pads to the longest row of the batch."""
        ln = self.tok_len[rows]
        if scaffold_prefix:
            ln = ln + self.sca_len[rows] + 1
        m = int(ln.max()) if len(rows) else 0
        return m, m + 2

    def collate(self, rows, model_type):
        """One batch as a dict of device tensors; `rows` = sequence of corpus row indices."""
        rows = np.asarray(rows, dtype=np.int64)
        sca = model_type in _SCAFFOLD_MODELS
        if sca and not self.use_scaffold:
            raise L.GctError(f"{model_type} needs a corpus built with use_scaffold=True")
        dev = self._dev["device"] if self._dev else None
        if dev is None:
            raise L.GctError("TokenisedCorpus: call .to(device) first")
        B = len(rows)
        S, T = self.batch_shape(rows, sca)
        rows_d = torch.from_numpy(rows).to(dev, non_blocking=True)
        src = torch.empty((B, S), dtype=torch.int64, device=dev)
        trg = torch.empty((B, T), dtype=torch.int64, device=dev)
        nc = len(self.property_list)
        want_c = nc > 0 and model_type != "vaetf"
        ec = torch.empty((B, nc), dtype=torch.float32, device=dev) if want_c else None
        dc = torch.empty((B, nc), dtype=torch.float32, device=dev) if want_c else None
        st, tt = self.SRC, self.TRG
        cs = self._struct()
        L.check(L.lib().gct_collate(C.byref(cs), rows_d.data_ptr(), B, S, T, _stoi(st, "<pad>"), _stoi(tt, "<pad>"), _stoi(tt, "<sos>"),
                                    _stoi(tt, "<eos>"), _stoi(st, "<sep>") if sca else -1, _stoi(tt, "<sep>") if sca else -1,
                                    src.data_ptr(), trg.data_ptr(), None if ec is None else ec.data_ptr(),
                                    None if dc is None else dc.data_ptr(), L.stream_ptr()), "gct_collate")
        out = {"src": src, "trg": trg}
        if want_c:
            out["econds"], out["dconds"] = ec, dc
        return out


class DeviceDataLoader:
    """Iterates batches like ``DataLoader(dataset, batch_size, drop_last=False, sampler=..., shuffle=...)`` of the
    reference, but every batch is assembled on the device.  ``sampler`` is one of torch's index samplers."""

    def __init__(self, corpus: TokenisedCorpus, model_type, batch_size, sampler):
        self.corpus, self.model_type, self.batch_size, self.sampler = corpus, model_type, int(batch_size), sampler

    def __len__(self):
        return (len(self.sampler) + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        # torch's DataLoader iterator draws its worker base seed from the global generator before the sampler draws its
        # permutation seed (torch/utils/data/dataloader.py, _BaseDataLoaderIter.__init__, same in the reference's torch 1.10):
        # consume the same number so that a seeded run visits the rows in the reference's order
        torch.empty((), dtype=torch.int64).random_()
        order = np.fromiter(iter(self.sampler), dtype=np.int64)
        for i in range(0, len(order), self.batch_size):
            yield self.corpus.collate(order[i:i + self.batch_size], self.model_type)


class _Rows:
    """Index-only stand-in for the dataset argument of torch's samplers."""

    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n


class DataloaderPreparation:
    """Same constructor / get_dataloader surface as Utils/dataset.py:292-335; `rank` doubles as the CUDA device index
    exactly as in the reference (collate_fn(..., device=rank))."""

    def __init__(self, rank, SRC, TRG, model_type, property_list, world_size=1, randomize_prob=False, use_scaffold=False):
        self.SRC, self.TRG, self.rank, self.world_size = SRC, TRG, rank, world_size
        self.model_type, self.property_list = model_type, property_list
        self.use_scaffold, self.randomize_prob = use_scaffold, randomize_prob

    def get_dataloader(self, dataframe, batch_size, is_train, include_mconds=False, shuffle=False, sampler=None):
        corpus = TokenisedCorpus(dataframe, self.property_list, self.SRC, self.TRG, self.use_scaffold, self.randomize_prob)
        corpus.to(torch.device("cuda", self.rank) if isinstance(self.rank, int) else self.rank)
        rows = _Rows(len(corpus))
        if self.world_size > 1:
            sampler = DistributedSampler(rows, self.world_size, self.rank if isinstance(self.rank, int) else 0,
                                         shuffle=bool(is_train), drop_last=False)
        elif sampler is None:
            sampler = RandomSampler(rows) if (is_train or shuffle) else SequentialSampler(rows)
        return DeviceDataLoader(corpus, self.model_type, batch_size, sampler)
