"""Generate tests/golden/attn_small.pt: the reference's secondary entry points run UNMODIFIED on CPU.

    python oracle/make_attn_golden.py          (build container only: needs /root/reference)

Recorded (small architecture, weights from tests/golden/fill.py, eval mode):
  * Vaetf(get_attn=True).forward -> the 8-tuple with the three attention lists        Model/vaetf.py:154-182, layers.py:24-25,60-61,69-70
  * VaetfSampling.get_attention_map(smiles)                                            Inference/sampling_tool.py:191-223
  * ScaVaeSampling.get_attention_map(smiles, scaffold)                                 Inference/sampling_tool.py:505-544
  * encode_smiles / encode_batch of the four samplers (mu / log_var; z depends on the CPU RNG draw)
                                                                                       Inference/sampling_tool.py:225-236,280-294,347-354,546-553
Same import shims as oracle/make_golden.py (none touches arithmetic).
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden import ARCH_SMALL, FakeField, FakeScaler, import_reference, make_batch  # noqa: E402
from fill import fill_state_dict  # noqa: E402


def build(ref, mt, nc, c2l, get_attn):
    cls = ref["Vaetf"] if mt == "vaetf" else ref["Cvaetf"]
    torch.manual_seed(0)
    m = cls(src_vocab=32, trg_vocab=32, dropout=0.1, nconds=nc, use_cond2dec=False, use_cond2lat=c2l, variational=True,
            get_attn=get_attn, **ARCH_SMALL)
    shapes = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    sd = fill_state_dict(shapes, seed=3)
    for k in list(sd):
        if sd[k] is None:
            sd[k] = m.state_dict()[k].clone()
    m.load_state_dict(sd)
    return m.eval(), shapes


def sampler_for(ref, m, mt, nc):
    kwargs = dict(top_k=None, latent_dim=ARCH_SMALL["latent_dim"], max_strlen=14, use_cond2dec=False, decode_algo="greedy",
                  n_jobs=1, toklen_data=np.array([6, 7, 8, 9, 10] * 100), cond_dim=nc, scaler=FakeScaler(), device="cpu",
                  SRC=FakeField(), TRG=FakeField())
    return ref["ST"].sampling_tool_dict[mt](m, kwargs)


def main():
    ref = import_reference()
    out = {"arch": ARCH_SMALL, "fill_seed": 3}
    smiles = ["CCOc1ccc", "c1ccncc1C(=O)N", "CC"]
    scaffolds = ["c1ccc", "c1ccncc1", "C"]
    econds = np.array([[1.0, 2.0, 3.0], [0.5, 1.5, 2.5], [2.0, 1.0, 0.0]])
    with torch.no_grad():
        # ---- Vaetf with get_attn=True: forward 8-tuple + get_attention_map
        m, shapes = build(ref, "vaetf", 0, False, True)
        batch = make_batch(3, 11, seed=21, nconds=0)
        trg_in = batch["trg"][:, :-1]
        sm = ref["M"].get_src_mask(batch["src"], 1)
        tm = ref["M"].get_trg_mask(trg_in, 1, False)
        torch.manual_seed(5)
        res = m(batch["src"], trg_in, sm, tm)
        assert len(res) == 8
        s = sampler_for(ref, m, "vaetf", 0)
        ea, d1, d2 = s.get_attention_map(smiles[0])
        # Vaetf.encode does not unpack the (x, attn) pair a get_attn=True encoder returns (vaetf.py:146-147): encode_smiles /
        # encode_batch only work on a get_attn=False model -- same weights
        m0, _ = build(ref, "vaetf", 0, False, False)
        s0 = sampler_for(ref, m0, "vaetf", 0)
        z, mu, lv = s0.encode_smiles(smiles)
        zb, mub, lvb = s0.encode_batch({"src": batch["src"].clone()})
        out["vaetf"] = dict(shapes=shapes, batch=batch, eps_seed=5, output_mol=res[1].clone(), mu=res[2].clone(), log_var=res[3].clone(),
                            z=res[4].clone(), enc_attn=[t.clone() for t in res[5]], dec_attn1=[t.clone() for t in res[6]],
                            dec_attn2=[t.clone() for t in res[7]], map_smiles=smiles[0], map_enc=[t.clone() for t in ea],
                            map_dec1=[t.clone() for t in d1], map_dec2=[t.clone() for t in d2], smiles=smiles, enc_mu=mu.clone(),
                            enc_log_var=lv.clone(), encb_mu=mub.clone(), encb_log_var=lvb.clone())
        # ---- scavaetf (Cvaetf, nconds=0) with get_attn=True: get_attention_map only (Cvaetf.forward does not unpack the
        #      encoder's 4-tuple, cvaetf.py:178, so the reference itself can only use get_attn through the sampler)
        m, shapes = build(ref, "scavaetf", 0, False, True)
        s = sampler_for(ref, m, "scavaetf", 0)
        with contextlib.redirect_stdout(io.StringIO()):
            ea, d1, d2 = s.get_attention_map(smiles[0], scaffolds[0])
        out["scavaetf_attn"] = dict(shapes=shapes, smiles=smiles[0], scaffold=scaffolds[0], map_enc=[t.clone() for t in ea],
                                    map_dec1=[t.clone() for t in d1], map_dec2=[t.clone() for t in d2])
        # ---- encode_smiles / encode_batch of the conditional samplers (get_attn=False models)
        m, shapes = build(ref, "scavaetf", 0, False, False)
        s = sampler_for(ref, m, "scavaetf", 0)
        z, mu, lv = s.encode_smiles(smiles, scaffolds)
        out["scavaetf"] = dict(shapes=shapes, smiles=smiles, scaffolds=scaffolds, enc_mu=mu.clone(), enc_log_var=lv.clone())
        m, shapes = build(ref, "pvaetf", 3, True, False)
        s = sampler_for(ref, m, "pvaetf", 3)
        z, mu, lv = s.encode_smiles(smiles, econds, transform=True)
        batch = make_batch(3, 11, seed=22, nconds=3)
        zb, mub, lvb = s.encode_batch({"src": batch["src"].clone(), "econds": batch["econds"].clone()}, transform=True)
        out["pvaetf"] = dict(shapes=shapes, smiles=smiles, econds=econds, enc_mu=mu.clone(), enc_log_var=lv.clone(), batch=batch,
                             encb_mu=mub.clone(), encb_log_var=lvb.clone())
        m, shapes = build(ref, "pscavaetf", 3, True, False)
        s = sampler_for(ref, m, "pscavaetf", 3)
        z, mu, lv = s.encode_smiles(smiles, scaffolds, econds, transform=True)
        out["pscavaetf"] = dict(shapes=shapes, smiles=smiles, scaffolds=scaffolds, econds=econds, enc_mu=mu.clone(),
                                enc_log_var=lv.clone())
    path = os.path.join(ROOT, "tests", "golden", "attn_small.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
