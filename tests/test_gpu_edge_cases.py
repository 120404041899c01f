"""Edge cases: real MOSES-like vocabulary sizes (not multiples of 8), batch of one, single-token sources,
sequences longer than one attention tile (SIMT fall-back in the bf16 tier), maximum decode length."""
import numpy as np
import pytest
import torch

from gpu_common import DEV
from helpers import FakeScaler, O, rel_err
from gct_plus_b200.Model import Cvaetf, Vaetf
from gct_plus_b200.Model.modules import get_src_mask, get_trg_mask
from gct_plus_b200.Train.trainer1 import FusedTrainer, loss_function

pytestmark = pytest.mark.gpu
ARCH = dict(N=2, d_model=128, dff=256, h=2, latent_dim=32)


def _oracle_cfg(mt, vs, vt, nc, c2l):
    return O.ModelCfg(model_type=mt, src_vocab=vs, trg_vocab=vt, N=2, d_model=128, dff=256, h=2, latent_dim=32, nconds=nc,
                      use_cond2lat=c2l)


def _batch(B, S, vs, seed, nc=0, min_len=1):
    g = torch.Generator().manual_seed(seed)
    src = torch.full((B, S), 1, dtype=torch.long)
    trg = torch.full((B, S + 2), 1, dtype=torch.long)
    for b in range(B):
        L = int(torch.randint(min_len, S + 1, (1,), generator=g)) if b else S
        toks = torch.randint(5, vs, (L,), generator=g)
        src[b, :L] = toks
        trg[b, 0], trg[b, L + 1] = 2, 3
        trg[b, 1:L + 1] = toks
    out = {"src": src, "trg": trg}
    if nc:
        out["econds"] = torch.randn(B, nc, generator=g)
        out["dconds"] = out["econds"].clone()
    return out


@pytest.mark.parametrize("dtype,tol", [("fp32", 1e-4), ("bf16", 1.5e-2)])
@pytest.mark.parametrize("B,S,vs,vt", [(1, 1, 25, 27), (3, 9, 26, 28), (2, 140, 25, 27), (5, 33, 30, 31)])
def test_odd_vocab_and_lengths_forward_backward(dtype, tol, B, S, vs, vt):
    torch.manual_seed(0)
    m = Cvaetf(vs, vt, dropout=0.0, nconds=3, use_cond2lat=True, compute_dtype=dtype, **ARCH).to(DEV).train()
    sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    batch = _batch(B, S, min(vs, vt), seed=B * 100 + S, nc=3)
    dev = {k: v.to(DEV) for k, v in batch.items()}
    eps = torch.randn(B, 3 + S, 32, generator=torch.Generator().manual_seed(1))
    trg_in = dev["trg"][:, :-1]
    logits, mu, lv, z, _ = m._run(dev["src"], trg_in, get_src_mask(dev["src"], 1, dev["econds"]),
                                  get_trg_mask(trg_in, 1, False, dev["dconds"]), dev["econds"], dev["dconds"], eps=eps.to(DEV))
    cfg = _oracle_cfg("pvaetf", vs, vt, 3, True)
    params = {k: v.clone().requires_grad_(not k.endswith("pe.pe")) for k, v in sd.items()}
    _, mol, mu_o, lv_o, _ = O.forward_propagation(params, cfg, batch, 1, eps)
    assert logits.shape == mol.shape
    assert rel_err(logits, mol) < tol and rel_err(mu, mu_o) < tol
    ys = dev["trg"][:, 1:].reshape(-1)
    loss = loss_function(0.5, None, logits, None, ys, mu, lv, False, 1)[0]
    loss_o = O.loss_function(0.5, None, mol, None, batch["trg"][:, 1:].reshape(-1), mu_o, lv_o, False, 1)[0]
    assert abs(float(loss) - float(loss_o)) < tol * abs(float(loss_o))
    if S > 120:
        # training backward is tiled for L <= 128 (one tcgen05 tile; the fp32 tier is smem-bound at ~104): longer
        # sequences run forward / sampling only and backward must fail loudly, never silently
        import gct_plus_b200._lib as L
        with pytest.raises(L.GctError, match="attention bwd"):
            loss.backward()
        return
    loss.backward()
    loss_o.backward()
    named = dict(m.named_parameters())
    gmax = max(float(p.grad.abs().max()) for p in params.values() if p.grad is not None)
    for k in ("out.weight", "out.bias", "decoder.embed.embed.weight", "encoder.embed_sentence.embed.weight",
              "encoder.layers.0.ff.linear_1.weight", "decoder.layers.1.attn_2.v_linear.weight", "decoder.fc_z.weight"):
        g, go = named[k].grad.cpu(), params[k].grad
        assert float((g - go).abs().max()) < (5e-4 if dtype == "fp32" else 6e-2) * max(float(go.abs().max()), 1e-3 * gmax), k


def test_decode_to_the_positional_table_limit():
    """max_strlen 190 (PE table has 200 rows): greedy tokens equal the oracle's un-cached loop (fp32 tier)."""
    from gct_plus_b200.Inference.sampling_tool import VaetfSampling
    from helpers import FakeField
    torch.manual_seed(3)
    m = Vaetf(25, 27, dropout=0.1, nconds=0, compute_dtype="fp32", **ARCH).to(DEV).eval()
    sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}

    class F27(FakeField):
        pass
    kwargs = dict(top_k=None, latent_dim=32, max_strlen=190, use_cond2dec=False, decode_algo="greedy", n_jobs=1,
                  toklen_data=np.array([5.0, 9.0]), cond_dim=0, scaler=FakeScaler(), device=DEV, SRC=F27(), TRG=F27(), sync_every=64)
    s = VaetfSampling(m, kwargs)
    g = torch.Generator().manual_seed(4)
    zs = torch.randn(2, 7, 32, generator=g)
    mask = torch.ones(2, 1, 7, dtype=torch.bool)
    mask[1, 0, 5:] = False
    ys0 = torch.full((2, 1), 2, dtype=torch.long)
    ys = s.decode(zs=zs.to(DEV), ys=ys0.to(DEV), src_mask=mask.to(DEV)).cpu()
    cfg = O.ModelCfg(model_type="vaetf", src_vocab=25, trg_vocab=27, N=2, d_model=128, dff=256, h=2, latent_dim=32)
    want = O.sampling_decode(sd, cfg, zs, ys0, mask, max_strlen=190)
    assert torch.equal(ys, want)


def test_fused_trainer_odd_vocab_runs_and_learns():
    """A few FusedTrainer steps on one batch with the real vocabulary sizes: the loss must go down."""
    torch.manual_seed(0)
    m = Cvaetf(26, 28, dropout=0.1, nconds=0, compute_dtype="bf16", **ARCH).to(DEV).train()
    tr = FusedTrainer(m, "scavaetf", pad_id=1, lr=3e-3, warmup=100000)
    tr.lr = 3e-3
    batch = {k: v.to(DEV) for k, v in _batch(16, 20, 26, seed=9).items()}
    losses = []
    for _ in range(12):
        tr.step(batch, 0.0)
        tr.lr = 3e-3
        losses.append(tr.read_losses()[0])
    assert all(np.isfinite(losses)) and losses[-1] < 0.7 * losses[0], losses
