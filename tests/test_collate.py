"""Input pipeline: oracle vs the reference's own collate_fn (golden), host-side corpus logic (CPU), and
the device batch assembly (gct_collate) bit-exact against both (GPU)."""
import os
import sys

import numpy as np
import pandas as pd
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import collate_oracle as CO  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "collate.pt")
CASES = ["vaetf", "pvaetf", "scavaetf", "pscavaetf"]


def _gold():
    g = torch.load(GOLD, weights_only=False)
    return g, pd.DataFrame(g["frame"])


def test_tokeniser_regex():
    tk = CO.MolTokenizer()
    assert tk("CCl(Br)[nH]c1%12=O") == ["C", "Cl", "(", "Br", ")", "[nH]", "c", "1", "%12", "=", "O"]
    tks = CO.MolTokenizer(add_sep=True)
    assert tks("c1ccccc1<sep>CCO") == ["c", "1", "c", "c", "c", "c", "c", "1", "<sep>", "C", "C", "O"]
    assert tks("CCO") == ["C", "C", "O"]
    assert tk("F/C=C\\F") == ["F", "/", "C", "=", "C", "\\", "F"]          # one literal backslash (Utils/field.py:16 is a non-raw string)
    # the reference's literal (a NON-raw string: its "\\\\" is the two-character regex for one backslash)
    ref = (r"(\[[^\]]+]|Br?|Cl?|N|O|S|P|F|I|b|c|n|o|s|p|\(|\)|\.|=|#|-|\+|" + "\\\\" +
           r"|\/|:|~|@|\?|>|\*|\$|\%[0-9]{2}|[0-9])")
    assert CO._PATTERN == ref
    assert CO._PATTERN == _gold()[0]["tokenizer_pattern"]          # recorded from the reference's source by make_collate_golden.py


def test_field_process_pads_to_the_longest_row():
    SRC, TRG = CO.smiles_fields(["C", "N"])
    s = SRC.process([["C", "N", "C"], ["N"], ["X"]])
    assert s.tolist() == [[2, 3, 2], [3, 1, 1], [0, 1, 1]]          # <unk>=0, <pad>=1; no init / eos for SRC
    t = TRG.process([["C", "N", "C"], ["N"]])
    assert t.tolist() == [[2, 4, 5, 4, 3], [2, 5, 3, 1, 1]]         # <sos>=2 ... <eos>=3 <pad>=1


@pytest.mark.parametrize("mt", CASES)
def test_oracle_collate_matches_reference_collate_fn(mt):
    g, df = _gold()
    c = g["cases"][mt]
    SRC, TRG = CO.smiles_fields(g["atoms"], c["add_sep"])
    got = list(CO.batches(df, c["order"], c["batch_size"], mt, SRC, TRG, c["property_list"], c["use_scaffold"]))
    assert len(got) == len(c["batches"])
    for a, b in zip(got, c["batches"]):
        assert set(a) == set(b)
        for k in b:
            assert a[k].dtype == b[k].dtype and torch.equal(a[k], b[k]), (mt, k)


def test_corpus_csr_matches_getitem():
    from gct_plus_b200.Utils.dataset import TokenisedCorpus
    g, df = _gold()
    SRC, TRG = CO.smiles_fields(g["atoms"], True)
    corpus = TokenisedCorpus(df, g["props"], SRC, TRG, use_scaffold=True)
    assert len(corpus) == len(df)
    for r in (0, 3, 5, len(df) - 1):
        it = CO.getitem(df.iloc[r], SRC, TRG, g["props"], True)
        a, b = corpus.tok_off[r], corpus.tok_off[r + 1]
        assert corpus.src_ids[a:b].tolist() == [SRC.vocab.stoi[t] for t in it["src"]]
        assert corpus.trg_ids[a:b].tolist() == [TRG.vocab.stoi[t] for t in it["trg"]]
        a, b = corpus.sca_off[r], corpus.sca_off[r + 1]
        assert corpus.sca_trg_ids[a:b].tolist() == [TRG.vocab.stoi[t] for t in it["trg_scaffold"]]
        assert np.allclose(corpus.econds[r], it["econds"]) and np.allclose(corpus.dconds[r], it["dconds"])
    rows = np.array([0, 3, 5])
    S, T = corpus.batch_shape(rows, True)
    assert S == max(corpus.tok_len[r] + corpus.sca_len[r] + 1 for r in rows) and T == S + 2
    with pytest.raises(Exception):
        TokenisedCorpus(df, [], SRC, TRG, randomize_prob=0.5)
    with pytest.raises(Exception):
        corpus.collate(rows, "pscavaetf")          # not uploaded: no CPU path


@pytest.mark.gpu
@pytest.mark.parametrize("mt", CASES)
def test_device_batches_equal_reference_batches(mt):
    from gct_plus_b200.Utils.dataset import DeviceDataLoader, TokenisedCorpus
    g, df = _gold()
    c = g["cases"][mt]
    SRC, TRG = CO.smiles_fields(g["atoms"], c["add_sep"])
    corpus = TokenisedCorpus(df, c["property_list"], SRC, TRG, use_scaffold=c["use_scaffold"]).to("cuda:0")
    got = list(DeviceDataLoader(corpus, mt, c["batch_size"], c["order"]))
    assert len(got) == len(c["batches"])
    for a, b in zip(got, c["batches"]):
        assert set(a) == set(b)
        for k in b:
            assert a[k].is_cuda and a[k].dtype == b[k].dtype and torch.equal(a[k].cpu(), b[k]), (mt, k)


@pytest.mark.gpu
def test_device_loader_follows_torch_samplers_and_feeds_forward_propagation():
    """world_size 2: each rank's DistributedSampler order, big ragged corpus, batches equal the oracle's; one batch goes
    through forward_propagation unchanged."""
    from torch.utils.data import DistributedSampler
    from gct_plus_b200.Utils.dataset import DataloaderPreparation, _Rows
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from make_collate_golden import ATOMS, synthetic_frame
    props = ["logP", "tPSA", "QED"]
    df = synthetic_frame(1001, seed=9)
    SRC, TRG = CO.smiles_fields(ATOMS, True)
    for rank in (0, 1):
        prep = DataloaderPreparation(0, SRC, TRG, "pscavaetf", props, world_size=2, use_scaffold=True)
        prep.rank = 0                               # one visible GPU: device index 0, sampler rank below
        dl = prep.get_dataloader(df, 64, is_train=True)
        dl.sampler = DistributedSampler(_Rows(len(df)), 2, rank, shuffle=True)
        dl.sampler.set_epoch(3)
        ref_sampler = DistributedSampler(_Rows(len(df)), 2, rank, shuffle=True)
        ref_sampler.set_epoch(3)
        want = list(CO.batches(df, list(ref_sampler), 64, "pscavaetf", SRC, TRG, props, True))
        got = list(dl)
        assert len(got) == len(want) == len(dl)
        for a, b in zip(got, want):
            for k in b:
                assert torch.equal(a[k].cpu(), b[k]), k
    from gpu_common import build_model
    from helpers import load_golden
    from gct_plus_b200.Model.forward_propagation1 import forward_propagation
    fx = load_golden("pscavaetf_small")
    m, _ = build_model(fx, "bf16")
    m.eval()
    batch = {k: (v % 32 if v.dtype == torch.int64 else v) for k, v in got[0].items()}
    out = forward_propagation["pscavaetf"](m, batch, 1, False)
    assert torch.isfinite(out[1]).all()


@pytest.mark.gpu
def test_reference_epoch_loop_on_the_device_loader(tmp_path):
    """train1.py-style run: DataloaderPreparation -> train_model (reference epoch loop, torch Adam through the autograd
    bridge, per-epoch CSVs + checkpoint) for two epochs on the device-assembled batches; the checkpoint reloads."""
    import argparse
    from gct_plus_b200.Model.build_model import get_model, load_state
    from gct_plus_b200.Train.trainer1 import train_model
    from gct_plus_b200.Utils.dataset import DataloaderPreparation
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from make_collate_golden import ATOMS, synthetic_frame
    props = ["logP", "tPSA", "QED"]
    SRC, TRG = CO.smiles_fields(ATOMS, True)
    df = synthetic_frame(96, seed=21)
    args = argparse.Namespace(N=2, d_model=128, d_ff=256, H=2, latent_dim=32, dropout=0.1, use_cond2dec=False, use_cond2lat=True,
                              variational=True, property_list=props, get_attn=False, model_type="pscavaetf", pad_id=1,
                              lr_scheduler="WarmUpDefault", lr_WarmUpSteps=50, start_epoch=1, num_epoch=2, use_KLA=True,
                              KLA_ini_beta=0.02, KLA_inc_beta=0.02, KLA_max_beta=1.0, KLA_beg_epoch=1,
                              model_folder=str(tmp_path))
    torch.manual_seed(0)
    model = get_model(args, len(SRC.vocab), len(TRG.vocab), 0).to("cuda:0")       # train1.py:108 does model.to(rank)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.98), eps=1e-9)
    prep = DataloaderPreparation(0, SRC, TRG, "pscavaetf", props, world_size=1, use_scaffold=True)
    train_loader = prep.get_dataloader(df, 32, is_train=True)
    valid_loader = prep.get_dataloader(df.iloc[:40].reset_index(drop=True), 32, is_train=False)
    train_model(args, model, opt, train_loader, valid_loader, 0, 1, None)
    t1 = pd.read_csv(tmp_path / "train_1.csv")
    t2 = pd.read_csv(tmp_path / "train_2.csv")
    assert len(t1) == len(train_loader) == 3 and np.isfinite(t1["LOSS"]).all() and np.isfinite(t2["LOSS"]).all()
    assert t2["RCE"].mean() < t1["RCE"].mean()                 # it learns
    assert os.path.exists(tmp_path / "valid_2.csv")
    fresh = get_model(args, len(SRC.vocab), len(TRG.vocab), 0).to("cuda:0")
    ck = torch.load(tmp_path / "model_2.pt", weights_only=False)
    fresh.load_state_dict(ck["model_state_dict"])
    for (k, a), (_, b) in zip(model.state_dict().items(), fresh.state_dict().items()):
        assert torch.equal(a, b), k


def test_device_loader_order_equals_torch_dataloader():
    """Host logic only (collate stubbed): the row order and batch boundaries are those of torch's DataLoader with the same
    sampler -- RandomSampler under a seed, DistributedSampler per rank and epoch -- including the ragged last batch."""
    from torch.utils.data import DataLoader, DistributedSampler, RandomSampler
    from gct_plus_b200.Utils.dataset import DeviceDataLoader, _Rows

    class _Corpus:
        def collate(self, rows, model_type):
            return list(map(int, rows))

    n, bs = 103, 16
    for make in (lambda: RandomSampler(_Rows(n)),
                 lambda: DistributedSampler(_Rows(n), 2, 1, shuffle=True),
                 lambda: DistributedSampler(_Rows(n), 4, 3, shuffle=False)):
        s1, s2 = make(), make()
        for s in (s1, s2):
            if hasattr(s, "set_epoch"):
                s.set_epoch(5)
        torch.manual_seed(123)
        got = list(DeviceDataLoader(_Corpus(), "vaetf", bs, s1))
        torch.manual_seed(123)
        want = [b.tolist() for b in DataLoader(list(range(n)), batch_size=bs, sampler=s2, drop_last=False)]
        assert got == want and len(DeviceDataLoader(_Corpus(), "vaetf", bs, s1)) == len(want)
