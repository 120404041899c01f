"""CPU oracle of the input pipeline.  TEST INFRASTRUCTURE ONLY: imported by tests/ and by
bench.py's cpu_baseline leg, never by the product (gct_plus_b200/Utils/dataset.py assembles batches on the device).

Restates, in plain Python / NumPy:
  * torchtext 0.6.0 ``Field.pad`` + ``Field.numericalize`` for a sequential, batch_first field without fixed length
    (third-party dependency of the reference, pinned ``torchtext==0.6.0`` in env.yml; absent from this image, so the
    published algorithm is restated: pad every example to the longest of the minibatch as
    [init_token] + tokens + [eos_token] + [pad_token]*(max_len - len), then map through ``vocab.stoi``);
  * the reference's collate functions (Model/collate_fn.py:4-124) and SmilesDataset.__getitem__ (Utils/dataset.py:270-289);
  * the regex tokeniser (Utils/field.py:8-33).
Pinning: ``process`` (torchtext) is restated, hence unpinned; the collate functions ARE pinned -- oracle/make_collate_golden.py
runs the reference's own Model/collate_fn.py (with this file's Field as the duck-typed SRC/TRG) and tests/test_collate.py
checks `collate` below against those recorded batches.
"""
import re

import numpy as np

_PATTERN = r"(\[[^\]]+]|Br?|Cl?|N|O|S|P|F|I|b|c|n|o|s|p|\(|\)|\.|=|#|-|\+|\\|\/|:|~|@|\?|>|\*|\$|\%[0-9]{2}|[0-9])"


class MolTokenizer:
    """Utils/field.py:8-33 (moltokenize)."""

    def __init__(self, add_sep=False):
        self.regex = re.compile(_PATTERN)
        self.add_sep = add_sep

    def _tokenize(self, s):
        return [t for t in self.regex.findall(s)]

    def __call__(self, s):
        if not self.add_sep:
            return [t for t in self._tokenize(s) if t != " "]
        res = re.split(r"(<sep>)", s)
        if len(res) == 1:
            return [t for t in self._tokenize(s) if t != " "]
        if len(res) == 3:
            return self._tokenize(res[0]) + ["<sep>"] + self._tokenize(res[2])
        return []


class Vocab:
    def __init__(self, itos):
        self.itos = list(itos)
        self._stoi = {t: i for i, t in enumerate(self.itos)}
        self.stoi = self                      # torchtext: defaultdict falling back to <unk> (index 0)

    def __getitem__(self, tok):
        return self._stoi.get(tok, 0)

    def get(self, tok, default=None):
        return self._stoi.get(tok, default)

    def __len__(self):
        return len(self.itos)


class Field:
    """torchtext 0.6.0 data.Field(tokenize=..., batch_first=True[, init_token, eos_token]) -- the parts the path touches."""
    batch_first = True
    pad_token = "<pad>"

    def __init__(self, tokenize, itos, init_token=None, eos_token=None):
        self.tokenize = tokenize
        self.vocab = Vocab(itos)
        self.init_token, self.eos_token = init_token, eos_token

    def pad(self, minibatch):
        max_len = max(len(x) for x in minibatch)
        out = []
        for x in minibatch:
            out.append(([] if self.init_token is None else [self.init_token]) + list(x[:max_len]) +
                       ([] if self.eos_token is None else [self.eos_token]) + [self.pad_token] * max(0, max_len - len(x)))
        return out

    def numericalize(self, arr):
        import torch
        return torch.tensor([[self.vocab.stoi[t] for t in ex] for ex in arr], dtype=torch.long)

    def process(self, batch, device=None):
        return self.numericalize(self.pad(batch))


def smiles_fields(tokens, add_sep=False):
    """Utils/field.py:47-53 with a vocabulary built from `tokens`; torchtext's special-token order:
    <unk>, <pad>, [init, eos,] then the remaining tokens."""
    tk = MolTokenizer(add_sep)
    extra = ["<sep>"] if add_sep else []
    SRC = Field(tk, ["<unk>", "<pad>"] + extra + list(tokens))
    TRG = Field(tk, ["<unk>", "<pad>", "<sos>", "<eos>"] + extra + list(tokens), init_token="<sos>", eos_token="<eos>")
    return SRC, TRG


def getitem(row, SRC, TRG, property_list, use_scaffold):
    """Utils/dataset.py:270-289 (randomize_prob = 0)."""
    item = {"src": SRC.tokenize(row["src"]), "trg": TRG.tokenize(row["src"])}
    if use_scaffold:
        item["src_scaffold"] = SRC.tokenize(row["src_scaffold"])
        item["trg_scaffold"] = TRG.tokenize(row["src_scaffold"])
    if len(property_list) > 0:
        item["econds"] = [row[f"src_{p}"] for p in property_list]
        item["dconds"] = [row[f"trg_{p}"] for p in property_list]
    return item


def collate(ins, model_type, SRC, TRG):
    """Model/collate_fn.py: vaetf :4-16, pvaetf :19-36, scavaetf / pscavaetf :104-124 (device = CPU)."""
    import torch
    outs = {}
    if model_type in ("scavaetf", "pscavaetf"):
        outs["src"] = SRC.process([b["src_scaffold"] + ["<sep>"] + b["src"] for b in ins])
        outs["trg"] = TRG.process([b["trg_scaffold"] + ["<sep>"] + b["trg"] for b in ins])
    else:
        outs["src"] = SRC.process([e["src"] for e in ins])
        outs["trg"] = TRG.process([e["trg"] for e in ins])
    if model_type != "vaetf":
        for p in ("econds", "dconds"):
            if p in ins[0]:
                outs[p] = torch.tensor([e[p] for e in ins], dtype=torch.float32)
    return outs


def batches(dataframe, order, batch_size, model_type, SRC, TRG, property_list, use_scaffold):
    """DataLoader(dataset, batch_size, drop_last=False, sampler=order, collate_fn=...) as a generator."""
    order = list(order)
    for i in range(0, len(order), batch_size):
        ins = [getitem(dataframe.iloc[int(r)], SRC, TRG, property_list, use_scaffold) for r in order[i:i + batch_size]]
        yield collate(ins, model_type, SRC, TRG)
