"""Whole-model parity: CUDA forward / backward / train step vs the oracle and the golden fixtures
(which hold the reference's own outputs).  Tolerances: fp32 tier 1e-4 relative, bf16 tier 1e-2
relative (north_star), both relative to the tensor's max magnitude."""
import pytest
import torch

from gpu_common import DEV, build_model
from helpers import CASE_NAMES, O, cfg_from_fixture, eps_for, load_golden, rel_err, sd_from_fixture
from gct_plus_b200.Model.forward_propagation1 import forward_propagation
from gct_plus_b200.Model.modules import get_src_mask, get_trg_mask
from gct_plus_b200.Train.trainer1 import FusedTrainer, loss_function

pytestmark = pytest.mark.gpu
TOL = {"fp32": 1e-4, "bf16": 1e-2}


def _to_dev(batch):
    return {k: v.to(DEV) for k, v in batch.items()}


def _run_with_eps(m, fx, batch, eps):
    nc = fx["nconds"]
    has_c = fx["model_type"] in ("pvaetf", "pscavaetf")
    trg_in = batch["trg"][:, :-1]
    sm = get_src_mask(batch["src"], 1, batch["econds"] if has_c else None)
    tm = get_trg_mask(trg_in, 1, fx.get("use_cond2dec", False), batch["dconds"] if has_c else None)
    return m._run(batch["src"], trg_in, sm, tm, batch.get("econds") if has_c else None, batch.get("dconds") if has_c else None,
                  eps=eps.to(DEV))


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("name", CASE_NAMES)
def test_forward_matches_reference_golden(name, dtype):
    fx = load_golden(name)
    m, _ = build_model(fx, dtype)
    m.eval()
    batch = _to_dev(fx["batch"])
    with torch.no_grad():
        logits, mu, lv, z, _ = _run_with_eps(m, fx, batch, eps_for(fx))
    nc = fx["nconds"]
    mol = logits[:, nc:, :] if fx.get("use_cond2dec") else logits
    tol = TOL[dtype]
    assert rel_err(mol, fx["output_mol"]) < tol
    assert rel_err(mu, fx["mu"]) < tol and rel_err(lv, fx["log_var"]) < tol and rel_err(z, fx["z"]) < tol


@pytest.mark.parametrize("name", CASE_NAMES)
def test_public_forward_propagation_and_loss(name):
    """Through the reference-shaped call surface (forward_propagation + loss_function), fp32 tier."""
    fx = load_golden(name)
    m, _ = build_model(fx, "fp32")
    m.eval()
    batch = _to_dev(fx["batch"])
    torch.manual_seed(fx["eps_seed"])
    prop, mol, mu, lv, z = forward_propagation[fx["model_type"]](m, batch, 1, fx.get("use_cond2dec", False))
    # eps came from the device RNG here, so compare what does not depend on it
    assert rel_err(mu, fx["mu"]) < 1e-4 and rel_err(lv, fx["log_var"]) < 1e-4
    nc = fx["nconds"]
    ys_cond = batch["dconds"].unsqueeze(2).view(-1, nc, 1) if nc else None
    loss, rce, rprop, kld = loss_function(fx["beta"], prop, mol, ys_cond, batch["trg"][:, 1:].reshape(-1), mu, lv,
                                          fx.get("use_cond2dec", False), 1)
    assert abs(float(kld) - fx["kld"]) < 1e-4 * abs(fx["kld"])
    assert mol.shape == fx["output_mol"].shape
    if prop is not None and fx["output_prop"] is not None:
        assert prop.shape == fx["output_prop"].shape


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("name", CASE_NAMES)
def test_backward_matches_reference_golden(name, dtype):
    """loss.backward() through the autograd bridge; gradients vs the reference's (golden) and the oracle's."""
    fx = load_golden(name)
    m, sd = build_model(fx, dtype, dropout=0.0)
    m.train()
    batch = _to_dev(fx["batch"])
    eps = eps_for(fx)
    logits, mu, lv, z, _ = _run_with_eps(m, fx, batch, eps)
    nc = fx["nconds"]
    c2d = fx.get("use_cond2dec", False)
    if c2d:
        prop = torch.nn.functional.linear(logits[:, :nc, :], m.prop_fc.weight, m.prop_fc.bias)
        mol = logits[:, nc:, :]
    else:
        prop, mol = None, logits
    ys_cond = batch["dconds"].unsqueeze(2).view(-1, nc, 1) if nc else None
    loss, rce, rprop, kld = loss_function(fx["beta"], prop, mol, ys_cond, batch["trg"][:, 1:].reshape(-1), mu, lv, c2d, 1)
    tol = TOL[dtype]
    assert abs(float(loss) - fx["loss"]) < tol * abs(fx["loss"])
    loss.backward()
    # oracle gradients (full tensors) on CPU
    cfg = cfg_from_fixture(fx)
    params = {k: v.clone().requires_grad_(not k.endswith("pe.pe")) for k, v in sd.items()}
    po, mo, muo, lvo, _ = O.forward_propagation(params, cfg, fx["batch"], 1, eps)
    ys_c = fx["batch"]["dconds"].unsqueeze(2).view(-1, nc, 1) if nc else None
    lo = O.loss_function(fx["beta"], po, mo, ys_c, fx["batch"]["trg"][:, 1:].reshape(-1), muo, lvo, c2d, 1)[0]
    lo.backward()
    gmax = max(float(p.grad.abs().max()) for p in params.values() if p.grad is not None)
    named = dict(m.named_parameters())
    worst = 0.0
    for k, p in params.items():
        if p.grad is None or k.endswith("k_linear.bias"):
            continue
        g = named[k].grad
        assert g is not None, k
        scale = max(float(p.grad.abs().max()), 1e-3 * gmax)
        err = float((g.cpu() - p.grad).abs().max()) / scale
        worst = max(worst, err)
        assert err < (5e-4 if dtype == "fp32" else 6e-2), (k, err)
    # the reference's own gradient summaries
    for k, s in fx["grads"].items():
        if k.endswith("k_linear.bias"):
            continue
        got = float(named[k].grad.double().abs().sum())
        assert abs(got - s["abssum"]) <= (2e-3 if dtype == "fp32" else 5e-2) * s["abssum"] + 1e-5 * s["numel"] + 1e-3, k


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_fused_trainer_three_steps_vs_reference(dtype):
    """FusedTrainer (no autograd, fused Adam, Noam LR applied after the step) against three optimiser steps
    of the reference recorded in tests/golden/train3.pt."""
    fx = load_golden("train3")
    fx = dict(fx, use_cond2lat=True)
    m, _ = build_model(fx, dtype, dropout=0.0)
    m.train()
    tr = FusedTrainer(m, fx["model_type"], pad_id=1, lr=1e-4, warmup=fx["warmup"])
    tol = 1e-4 if dtype == "fp32" else 1e-2
    for step, batch in enumerate(fx["batches"], start=1):
        B, S = batch["src"].shape
        torch.manual_seed(fx["eps_seeds"][step - 1])
        eps = torch.randn(B, fx["nconds"] + S, fx["arch"]["latent_dim"])
        tr.step(_to_dev(batch), fx["beta"], eps_noise=eps.to(DEV))
        loss, rce, kld = tr.read_losses()
        h = fx["hist"][step - 1]
        assert abs(loss - h["loss"]) < tol * abs(h["loss"]), (step, loss, h["loss"])
        assert abs(tr.lr - h["lr"]) < 1e-12
    if dtype == "fp32":
        sd = m.state_dict()
        for k, full in fx["final_full"].items():
            if k.endswith("pe.pe") or k.endswith("k_linear.bias"):
                continue
            assert rel_err(sd[k], full) < 2e-4, k


def test_backward_in_the_preactivation_form_of_the_ffn():
    """gct_set_ffn_saved_activation(1): the forward saves the FFN pre-activation and the backward epilogue evaluates
    gelu' (EPI_DGELU) instead of multiplying by the saved keep*gelu' -- same gradients as the default form."""
    import gct_plus_b200._lib as L
    try:
        L.lib().gct_set_ffn_saved_activation(1)
        test_backward_matches_reference_golden("pvaetf_full", "bf16")
    finally:
        L.lib().gct_set_ffn_saved_activation(0)


def test_dropout_is_statistically_right_and_backward_consistent():
    """Train mode: the dropout mask is regenerated in backward from (seed, site, index).  Check the keep rate
    through the PE dropout (the only one directly observable) and that two forwards with the same seed agree."""
    fx = load_golden("pscavaetf_small")
    m, _ = build_model(fx, "fp32", dropout=0.1)
    m.train()
    batch = _to_dev(fx["batch"])
    eps = eps_for(fx)
    m._step_seed = 1234
    a = _run_with_eps(m, fx, batch, eps)[0].detach().clone()
    m._step_seed = 1234
    b = _run_with_eps(m, fx, batch, eps)[0].detach().clone()
    assert torch.equal(a, b)
    m._step_seed = 99
    c = _run_with_eps(m, fx, batch, eps)[0].detach()
    assert not torch.equal(a, c)
    m.eval()
    with torch.no_grad():
        e = _run_with_eps(m, fx, batch, eps)[0]
    assert rel_err(e, fx["output_mol"]) < 1e-4


def test_full_size_properties_cfg1_shape():
    """BASELINE cfg 1 shape (vaetf, B=128, S=78, T=79) on the GPU: bf16 vs fp32 tier agree to 1e-2, the loss is finite,
    and padding rows do not influence non-padded outputs (mask property)."""
    torch.manual_seed(0)
    from gct_plus_b200.Model import Vaetf
    m = Vaetf(32, 32, N=6, d_model=512, dff=2048, h=8, latent_dim=128, dropout=0.1, nconds=0).to(DEV).eval()
    g = torch.Generator().manual_seed(1)
    B, S = 128, 78
    toks = torch.randint(5, 32, (B, S), generator=g)
    lens = torch.randint(20, S + 1, (B,), generator=g)
    src = toks.clone()
    trg = torch.full((B, S + 2), 1, dtype=torch.long)
    for b in range(B):
        src[b, lens[b]:] = 1
        trg[b, 0] = 2
        trg[b, 1:lens[b] + 1] = toks[b, :lens[b]]
        trg[b, lens[b] + 1] = 3
    src, trg = src.to(DEV), trg.to(DEV)
    trg_in = trg[:, :-1]
    eps = torch.randn(B, S, 128, generator=g).to(DEV)
    outs = {}
    with torch.no_grad():
        for dt in ("fp32", "bf16"):
            m.set_compute_dtype(dt)
            outs[dt] = m._run(src, trg_in, get_src_mask(src, 1), get_trg_mask(trg_in, 1, False), None, None, eps=eps)
        assert rel_err(outs["bf16"][0], outs["fp32"][0]) < 1e-2
        assert rel_err(outs["bf16"][1], outs["fp32"][1]) < 1e-2
        # changing tokens behind the padding must not change anything (they are masked as keys, and their own rows are ignored)
        m.set_compute_dtype("fp32")
        src2 = src.clone()
        b0 = int((lens < S).nonzero()[0])
        assert src2[b0, -1] == 1
        logits_a = outs["fp32"][0]
        mu_a = outs["fp32"][1]
        src2[b0, -1] = 1   # still pad: identical
        again = m._run(src2, trg_in, get_src_mask(src2, 1), get_trg_mask(trg_in, 1, False), None, None, eps=eps)
        assert torch.equal(again[0], logits_a) and torch.equal(again[1], mu_a)
    ys = trg[:, 1:].reshape(-1)
    loss, rce, _, kld = loss_function(1.0, None, outs["fp32"][0], None, ys, outs["fp32"][1], outs["fp32"][2], False, 1)
    assert torch.isfinite(loss) and float(rce) > 0 and float(kld) > 0


def test_fused_trainer_checkpoint_round_trip(tmp_path):
    """FusedTrainer.state_dict() is a torch.optim.Adam state_dict: save_checkpoint writes the reference's checkpoint layout,
    torch.optim.Adam loads it, and a FusedTrainer resumed from it continues like the uninterrupted run."""
    import argparse
    from gct_plus_b200.Train.trainer1 import save_checkpoint
    fx = load_golden("pvaetf_plain_small")
    batch = _to_dev(fx["batch"])
    eps = eps_for(fx).to(DEV)

    def fresh():
        m, _ = build_model(fx, "fp32", dropout=0.0)
        m.train()
        return m, FusedTrainer(m, fx["model_type"], pad_id=1, lr=1e-3, warmup=50)

    m1, t1 = fresh()
    for _ in range(2):
        t1.step(batch, 0.3, eps_noise=eps)
    a = fx["arch"]
    args = argparse.Namespace(N=a["N"], d_model=a["d_model"], d_ff=a["dff"], H=a["h"], latent_dim=a["latent_dim"], dropout=0.0,
                              use_cond2dec=False, use_cond2lat=fx.get("use_cond2lat", False), variational=True,
                              property_list=["a"] * fx["nconds"])
    path = tmp_path / "model_1.pt"
    save_checkpoint(args, m1, t1, path)
    ck = torch.load(path, weights_only=False)
    assert set(ck) == {"model_state_dict", "opt_state_dict", "model_params"}
    adam = torch.optim.Adam(m1.parameters(), lr=1e-4, betas=(0.9, 0.98), eps=1e-9)
    adam.load_state_dict(ck["opt_state_dict"])                   # torch accepts the layout
    p0 = next(iter(m1.parameters()))
    assert float(adam.state[p0]["step"]) == 2.0 and adam.param_groups[0]["lr"] == pytest.approx(t1.lr)
    m2, t2 = fresh()
    m2.load_state_dict(ck["model_state_dict"])
    t2.load_state_dict(ck["opt_state_dict"])
    t1.step(batch, 0.3, eps_noise=eps)
    t2.step(batch, 0.3, eps_noise=eps)
    assert t2.step_count == 3
    for (k, x), (_, y) in zip(m1.state_dict().items(), m2.state_dict().items()):
        if k.endswith("k_linear.bias"):
            continue          # mathematically zero gradient: rounding noise normalised by Adam (see test_fused_trainer...)
        assert rel_err(x, y) < 1e-4, k


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_fused_trainer_property_head_matches_reference(dtype):
    """use_cond2dec through FusedTrainer (property rows in front of the target rows, prop_fc + MSE(sum) from
    gct_prop_head_fwd_bwd): loss terms against the reference's recorded values, every gradient -- prop_fc's included --
    against the oracle's autograd.  cvaetf.py:103-105,184-186; trainer1.py:24-26."""
    fx = load_golden("pvaetf_c2d_small")
    m, sd = build_model(fx, dtype, dropout=0.0)
    m.train()
    tr = FusedTrainer(m, "pvaetf", pad_id=1)
    assert tr.use_cond2dec
    eps = eps_for(fx)
    tol = TOL[dtype]
    # validation pass first (train=False: loss terms only, parameters untouched), then the optimiser step
    tr.step(_to_dev(fx["batch"]), fx["beta"], eps_noise=eps.to(DEV), train=False)
    l0 = tr.read_losses(with_prop=True)
    assert abs(l0[0] - fx["loss"]) < tol * abs(fx["loss"]) and abs(l0[2] - fx["rprop"]) < tol * abs(fx["rprop"])
    tr.step(_to_dev(fx["batch"]), fx["beta"], eps_noise=eps.to(DEV))
    loss, rce, rprop, kld = tr.read_losses(with_prop=True)
    assert abs(loss - fx["loss"]) < tol * abs(fx["loss"]) and abs(rce - fx["rce"]) < tol * abs(fx["rce"])
    assert abs(rprop - fx["rprop"]) < tol * abs(fx["rprop"]) and abs(kld - fx["kld"]) < tol * abs(fx["kld"])
    cfg = cfg_from_fixture(fx)
    nc = fx["nconds"]
    params = {k: v.clone().requires_grad_(not k.endswith("pe.pe")) for k, v in sd.items()}
    po, mo, muo, lvo, _ = O.forward_propagation(params, cfg, fx["batch"], 1, eps)
    ys_c = fx["batch"]["dconds"].unsqueeze(2).view(-1, nc, 1)
    O.loss_function(fx["beta"], po, mo, ys_c, fx["batch"]["trg"][:, 1:].reshape(-1), muo, lvo, True, 1)[0].backward()
    gmax = max(float(p.grad.abs().max()) for p in params.values() if p.grad is not None)
    got = dict(zip([n for n, _ in m.named_parameters()], m.grad_views(tr.grads)))
    for k, p in params.items():
        if p.grad is None or k.endswith("k_linear.bias"):
            continue
        scale = max(float(p.grad.abs().max()), 1e-3 * gmax)
        err = float((got[k].cpu() - p.grad).abs().max()) / scale
        assert err < (5e-4 if dtype == "fp32" else 6e-2), (k, err)
    assert float(got["prop_fc.weight"].abs().max()) > 0


@pytest.mark.parametrize("fusion", [0, 2])
@pytest.mark.parametrize("name", ["vaetf_full", "pvaetf_full"])
def test_forward_backward_with_and_without_the_fused_residual_norm(name, fusion):
    """The bf16 forward at d_model = 512 with the residual projections fused with their Norm forced on (2) and off (0): both
    against the reference's recorded outputs; gradients of the two settings against each other (the backward reads the saved
    x / normalised activations whichever kernel wrote them)."""
    import gct_plus_b200._lib as L
    fx = load_golden(name)
    lib = L.lib()
    lib.gct_set_rownorm_fusion(fusion)
    try:
        m, _ = build_model(fx, "bf16", dropout=0.0)
        m.train()
        batch = _to_dev(fx["batch"])
        logits, mu, lv, z, _ = _run_with_eps(m, fx, batch, eps_for(fx))
        assert rel_err(logits, fx["output_mol"]) < 1e-2 and rel_err(mu, fx["mu"]) < 1e-2 and rel_err(z, fx["z"]) < 1e-2
        loss = loss_function(fx["beta"], None, logits, None, batch["trg"][:, 1:].reshape(-1), mu, lv, False, 1)[0]
        assert abs(float(loss) - fx["loss"]) < 1e-2 * abs(fx["loss"])
        loss.backward()
        for k, sg in fx["grads"].items():
            if k.endswith("k_linear.bias"):
                continue
            got = float(dict(m.named_parameters())[k].grad.double().abs().sum())
            assert abs(got - sg["abssum"]) <= 5e-2 * sg["abssum"] + 1e-5 * sg["numel"] + 1e-3, (k, fusion)
    finally:
        lib.gct_set_rownorm_fusion(0)
