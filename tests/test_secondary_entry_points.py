"""SURVEY.md 8f rank 4 -- the secondary entry points, against outputs of the UNMODIFIED reference recorded in
tests/golden/attn_small.pt (oracle/make_attn_golden.py): `get_attn=True` attention-probability outputs
(Model/layers.py:24-25,60-61,69-70; Model/vaetf.py:154-182), `get_attention_map` (Inference/sampling_tool.py:191-223,
505-544) and `encode_smiles` / `encode_batch` of the four samplers (:225-236, 280-294, 347-354, 546-553).
CPU part: the oracle restatement is pinned on the same fixture.  GPU part: the CUDA path through the public call surface."""
import numpy as np
import pytest
import torch

from helpers import FakeField, FakeScaler, O, load_golden, rel_err
from fill import fill_state_dict

TOL = {"fp32": 1e-4, "bf16": 1e-2}


def _sd(fx, key):
    sd = fill_state_dict(fx[key]["shapes"], seed=fx["fill_seed"])
    for k in sd:
        if sd[k] is None:
            sd[k] = O.positional_table(200, fx["arch"]["d_model"]).unsqueeze(0)
    return sd


def _cfg(fx, mt, nc=0, c2l=False):
    a = fx["arch"]
    return O.ModelCfg(model_type=mt, src_vocab=32, trg_vocab=32, N=a["N"], d_model=a["d_model"], dff=a["dff"], h=a["h"],
                      latent_dim=a["latent_dim"], nconds=nc, use_cond2lat=c2l)


def _eps(fx):
    v = fx["vaetf"]
    B, S = v["batch"]["src"].shape
    torch.manual_seed(v["eps_seed"])
    return torch.randn(B, S, fx["arch"]["latent_dim"])


def test_oracle_attention_outputs_match_reference():
    fx = load_golden("attn_small")
    v, sd, cfg = fx["vaetf"], _sd(fx, "vaetf"), _cfg(fx, "vaetf")
    src, trg_in = v["batch"]["src"], v["batch"]["trg"][:, :-1]
    sm = O.src_mask(src, 1)
    x, ea = O.encoder_trunk(sd, cfg, src, sm, None, want_attn=True)
    z, mu, lv = O.latent_heads(sd, cfg, x, _eps(fx))
    assert rel_err(mu, v["mu"]) < 1e-5 and rel_err(z, v["z"]) < 1e-5
    y, d1, d2 = O.decoder_trunk(sd, cfg, trg_in, z, sm, O.trg_mask(trg_in, 1), want_attn=True)
    assert rel_err(O.linear(y, sd, "out"), v["output_mol"]) < 1e-5
    for got, want in ((ea, v["enc_attn"]), (d1, v["dec_attn1"]), (d2, v["dec_attn2"])):
        assert len(got) == len(want) == cfg.N
        for a, b in zip(got, want):
            assert a.shape == b.shape and rel_err(a, b) < 1e-5


# ------------------------------------------------------------------------------------------ GPU
def _model(fx, key, mt, dtype, nc=0, c2l=False, get_attn=False):
    from gct_plus_b200.Model import Cvaetf, Vaetf
    a = fx["arch"]
    cls = Vaetf if mt == "vaetf" else Cvaetf
    m = cls(32, 32, N=a["N"], d_model=a["d_model"], dff=a["dff"], h=a["h"], latent_dim=a["latent_dim"], dropout=0.1, nconds=nc,
            use_cond2lat=c2l, get_attn=get_attn, compute_dtype=dtype)
    m.load_state_dict(_sd(fx, key))
    return m.to("cuda:0").eval()


def _sampler(m, mt, nc, fx):
    from gct_plus_b200.Inference.sampling_tool import sampling_tool_dict
    kwargs = dict(top_k=None, latent_dim=fx["arch"]["latent_dim"], max_strlen=14, use_cond2dec=False, decode_algo="greedy", n_jobs=1,
                  toklen_data=np.array([6, 7, 8, 9, 10] * 100), cond_dim=nc, scaler=FakeScaler(), device="cuda:0", SRC=FakeField(),
                  TRG=FakeField())
    return sampling_tool_dict[mt](m, kwargs)


def _cmp_lists(got, want, tol):
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert tuple(a.shape) == tuple(b.shape)
        assert rel_err(a, b) < tol


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_get_attn_forward_outputs_match_reference(dtype):
    from gct_plus_b200.Model.modules import get_src_mask, get_trg_mask
    fx = load_golden("attn_small")
    v = fx["vaetf"]
    m = _model(fx, "vaetf", "vaetf", dtype, get_attn=True)
    src, trg_in = v["batch"]["src"].cuda(), v["batch"]["trg"][:, :-1].cuda()
    sm, tm = get_src_mask(src, 1), get_trg_mask(trg_in, 1, False)
    tol = TOL[dtype]
    with torch.no_grad():
        res = m(src, trg_in, sm, tm)                       # the public call: 8-tuple (Model/vaetf.py:178-180)
        assert len(res) == 8
        _cmp_lists(res[5], v["enc_attn"], tol)             # encoder probabilities do not depend on the eps draw
        assert rel_err(res[2], v["mu"]) < tol and rel_err(res[3], v["log_var"]) < tol
        for lst, ref in ((res[6], v["dec_attn1"]), (res[7], v["dec_attn2"])):
            assert len(lst) == len(ref) and all(tuple(a.shape) == tuple(b.shape) for a, b in zip(lst, ref))
        # with the reference's eps: everything, decoder probabilities included
        logits, mu, lv, z, att = m._run(src, trg_in, sm, tm, None, None, eps=_eps(fx).cuda(), want_attn=True)
    assert rel_err(logits, v["output_mol"]) < tol and rel_err(z, v["z"]) < tol
    _cmp_lists(list(att[0]), v["enc_attn"], tol)
    _cmp_lists(list(att[1]), v["dec_attn1"], tol)
    _cmp_lists(list(att[2]), v["dec_attn2"], tol)
    for a in att[1]:                                       # rows are probability distributions; future keys get exactly 0
        assert float((a.sum(-1) - 1).abs().max()) < 1e-3
        assert float(torch.triu(a, diagonal=1).abs().max()) == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_get_attention_map_matches_reference(dtype):
    fx = load_golden("attn_small")
    tol = TOL[dtype]
    v = fx["vaetf"]
    s = _sampler(_model(fx, "vaetf", "vaetf", dtype, get_attn=True), "vaetf", 0, fx)
    ea, d1, d2 = s.get_attention_map(v["map_smiles"])
    _cmp_lists(ea, v["map_enc"], tol)
    _cmp_lists(d1, v["map_dec1"], tol)
    _cmp_lists(d2, v["map_dec2"], tol)
    sv = fx["scavaetf_attn"]
    s = _sampler(_model(fx, "scavaetf_attn", "scavaetf", dtype, get_attn=True), "scavaetf", 0, fx)
    ea, d1, d2 = s.get_attention_map(sv["smiles"], sv["scaffold"])
    _cmp_lists(ea, sv["map_enc"], tol)
    _cmp_lists(d1, sv["map_dec1"], tol)
    _cmp_lists(d2, sv["map_dec2"], tol)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_encode_smiles_and_encode_batch_match_reference(dtype):
    """mu / log_var of the four samplers' encode_smiles (and encode_batch) equal the reference's; z = mu + eps*std is drawn
    from the device RNG here (the reference draws on the CPU), so it is checked for shape and for following N(mu, std)."""
    fx = load_golden("attn_small")
    tol = TOL[dtype]
    v = fx["vaetf"]
    s = _sampler(_model(fx, "vaetf", "vaetf", dtype), "vaetf", 0, fx)
    z, mu, lv = s.encode_smiles(v["smiles"])
    assert rel_err(mu, v["enc_mu"]) < tol and rel_err(lv, v["enc_log_var"]) < tol and z.shape == mu.shape
    zs = torch.stack([s.encode_smiles(v["smiles"])[0] for _ in range(64)])         # 64 draws of z
    std = torch.exp(0.5 * lv)
    assert float(((zs.mean(0) - mu) / std).abs().max()) < 0.9                       # |mean error| < 0.9 sigma over 64 draws (7 sigma/sqrt(64))
    z, mu, lv = s.encode_batch({"src": v["batch"]["src"].clone()})
    assert rel_err(mu, v["encb_mu"]) < tol and rel_err(lv, v["encb_log_var"]) < tol
    c = fx["scavaetf"]
    s = _sampler(_model(fx, "scavaetf", "scavaetf", dtype), "scavaetf", 0, fx)
    z, mu, lv = s.encode_smiles(c["smiles"], c["scaffolds"])
    assert rel_err(mu, c["enc_mu"]) < tol and rel_err(lv, c["enc_log_var"]) < tol
    p = fx["pvaetf"]
    s = _sampler(_model(fx, "pvaetf", "pvaetf", dtype, nc=3, c2l=True), "pvaetf", 3, fx)
    z, mu, lv = s.encode_smiles(p["smiles"], p["econds"], transform=True)
    assert rel_err(mu, p["enc_mu"]) < tol and rel_err(lv, p["enc_log_var"]) < tol
    z, mu, lv = s.encode_batch({"src": p["batch"]["src"].clone(), "econds": p["batch"]["econds"].clone()}, transform=True)
    assert rel_err(mu, p["encb_mu"]) < tol and rel_err(lv, p["encb_log_var"]) < tol
    q = fx["pscavaetf"]
    s = _sampler(_model(fx, "pscavaetf", "pscavaetf", dtype, nc=3, c2l=True), "pscavaetf", 3, fx)
    z, mu, lv = s.encode_smiles(q["smiles"], q["scaffolds"], q["econds"], transform=True)
    assert rel_err(mu, q["enc_mu"]) < tol and rel_err(lv, q["enc_log_var"]) < tol
