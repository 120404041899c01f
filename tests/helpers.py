"""Shared test helpers: golden loading, oracle state_dict construction, fake Fields."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, GOLDEN)
sys.path.insert(0, ROOT)

from fill import fill_state_dict  # noqa: E402
from oracle import gct_oracle as O  # noqa: E402

CASE_NAMES = ["vaetf_full", "pvaetf_full", "scavaetf_small", "pscavaetf_small", "pvaetf_c2d_small",
              "pvaetf_plain_small"]


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def cfg_from_fixture(fx) -> O.ModelCfg:
    a = fx["arch"]
    return O.ModelCfg(model_type=fx["model_type"], src_vocab=32, trg_vocab=32, N=a["N"], d_model=a["d_model"],
                      dff=a["dff"], h=a["h"], latent_dim=a["latent_dim"], nconds=fx["nconds"],
                      use_cond2dec=fx.get("use_cond2dec", False), use_cond2lat=fx.get("use_cond2lat", True))


def sd_from_fixture(fx):
    sd = fill_state_dict(fx["shapes"], seed=fx["fill_seed"])
    d = fx["arch"]["d_model"]
    for k in sd:
        if sd[k] is None:
            sd[k] = O.positional_table(200, d).unsqueeze(0)
    return sd


def eps_for(fx, seed_key="eps_seed"):
    B, S = fx["batch"]["src"].shape
    torch.manual_seed(fx[seed_key])
    return torch.randn(B, fx["nconds"] + S, fx["arch"]["latent_dim"])


ITOS = ["<unk>", "<pad>", "<sos>", "<eos>", "<sep>"] + list("CcNnOoSsFIBrl()[]=#123456+-H@/")[:27]


class FakeVocab:
    def __init__(self):
        self.itos = ITOS
        self.stoi = {t: i for i, t in enumerate(ITOS)}

    def __len__(self):
        return len(self.itos)


class FakeField:
    batch_first = True

    def __init__(self):
        self.vocab = FakeVocab()

    def tokenize(self, smi):
        out, i = [], 0
        while i < len(smi):
            if smi.startswith("<sep>", i):
                out.append("<sep>")
                i += 5
            else:
                out.append(smi[i])
                i += 1
        return out

    def process(self, batch):
        L = max(len(x) for x in batch)
        ids = [[self.vocab.stoi[t] for t in x] + [1] * (L - len(x)) for x in batch]
        return torch.tensor(ids, dtype=torch.long)


class FakeScaler:
    def transform(self, x):
        return (np.asarray(x, dtype=np.float64) - 1.5) / 2.0


def rel_err(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))
