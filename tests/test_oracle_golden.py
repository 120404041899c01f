"""Pins oracle/gct_oracle.py against outputs of the reference's own modules (tests/golden)."""
import numpy as np
import pytest
import torch

from helpers import CASE_NAMES, O, cfg_from_fixture, eps_for, load_golden, rel_err, sd_from_fixture

TOL = 2e-5   # fp32 CPU, same BLAS; differences come only from op ordering


def test_pe_and_masks_and_norm():
    m = load_golden("misc")
    pe = O.positional_table(200, 512)
    assert torch.allclose(pe[[0, 1, 2, 57, 199]], m["pe512_rows"], atol=1e-6)
    assert abs(float(pe.double().sum()) - m["pe512_sum"]) < 1e-2
    assert torch.allclose(O.positional_table(200, 128)[[0, 1, 99]], m["pe128_rows"], atol=1e-6)
    t = m["mask_in"]
    assert torch.equal(O.src_mask(t, 1), m["src_mask"])
    assert torch.equal(O.src_mask(t, 1, 3), m["src_mask_c"])
    assert torch.equal(O.trg_mask(t, 1, False), m["trg_mask"] != 0)
    assert torch.equal(O.trg_mask(t, 1, True, 3), m["trg_mask_c2d"] != 0)
    y = O.norm(m["norm_x"], m["norm_alpha"], m["norm_bias"])
    assert torch.allclose(y, m["norm_y"], atol=2e-6)


def test_toklen_and_kla():
    m = load_golden("misc")
    np.random.seed(11)
    d = m["toklen_data"]
    out = O.toklen_from_distribution(d, 64, int(d.max() - d.min()))
    assert np.array_equal(out, m["toklen_out"])
    assert [O.kl_annealer(e, 0.02, 0.02, 1) for e in range(1, 6)] == m["kla"]


@pytest.mark.parametrize("name", CASE_NAMES)
def test_forward_loss_grads(name):
    fx = load_golden(name)
    cfg, sd = cfg_from_fixture(fx), sd_from_fixture(fx)
    params = {k: v.clone().requires_grad_(not k.endswith("pe.pe")) for k, v in sd.items()}
    eps = eps_for(fx)
    prop, mol, mu, lv, z = O.forward_propagation(params, cfg, fx["batch"], 1, eps)
    assert rel_err(mol, fx["output_mol"]) < TOL
    assert rel_err(mu, fx["mu"]) < TOL and rel_err(lv, fx["log_var"]) < TOL and rel_err(z, fx["z"]) < TOL
    if fx["output_prop"] is not None and fx["output_prop"].numel():
        assert rel_err(prop, fx["output_prop"]) < TOL
    nc = fx["nconds"]
    ys_cond = fx["batch"]["dconds"].unsqueeze(2).view(-1, nc, 1) if nc else None
    ys_mol = fx["batch"]["trg"][:, 1:].contiguous().view(-1)
    loss, rce, rprop, kld = O.loss_function(fx["beta"], prop, mol, ys_cond, ys_mol, mu, lv, cfg.use_cond2dec, 1)
    for got, key in ((loss, "loss"), (rce, "rce"), (kld, "kld")):
        assert abs(float(got.detach()) - fx[key]) <= 2e-5 * abs(fx[key]) + 1e-4
    loss.backward()
    for k, s in fx["grads"].items():
        g = params[k].grad
        assert g is not None, k
        assert abs(float(g.double().abs().sum()) - s["abssum"]) <= 1e-3 * s["abssum"] + 1e-6 * s["numel"] + 1e-4, k
    for k, gfull in fx["grad_full"].items():
        assert float((params[k].grad - gfull).abs().max()) <= 2e-4 * float(gfull.abs().max()) + 2e-5, k


@pytest.mark.parametrize("name", CASE_NAMES)
def test_encode_and_greedy_decode(name):
    fx = load_golden(name)
    cfg, sd = cfg_from_fixture(fx), sd_from_fixture(fx)
    nc = fx["nconds"]
    b = fx["batch"]
    B, S = b["src"].shape
    torch.manual_seed(fx["encode"]["seed"])
    eps = torch.randn(B, nc + S, cfg.latent_dim)
    z, mu, lv = O.encode(sd, cfg, b["src"], O.src_mask(b["src"], 1, nc), b.get("econds"), eps)
    assert rel_err(z, fx["encode"]["z"]) < TOL and rel_err(mu, fx["encode"]["mu"]) < TOL
    d = fx["decode"]
    ys = O.sampling_decode(sd, cfg, d["zs"], d["ys0"], d["src_mask"], d.get("dconds"),
                           max_strlen=d["max_strlen"], algo="greedy")
    assert torch.equal(ys, d["ys"])


def test_adam_noam_three_steps():
    fx = load_golden("train3")
    fx = dict(fx, use_cond2lat=True)
    cfg = cfg_from_fixture(fx)
    sd = sd_from_fixture(fx)
    params = {k: v.clone().requires_grad_(not k.endswith("pe.pe")) for k, v in sd.items()}
    state = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in params.items() if v.requires_grad}
    lr = 1e-4
    for step, batch in enumerate(fx["batches"], start=1):
        B, S = batch["src"].shape
        torch.manual_seed(fx["eps_seeds"][step - 1])
        eps = torch.randn(B, cfg.nconds + S, cfg.latent_dim)
        prop, mol, mu, lv, _ = O.forward_propagation(params, cfg, batch, 1, eps)
        ys_cond = batch["dconds"].unsqueeze(2).view(-1, cfg.nconds, 1)
        loss, rce, _, kld = O.loss_function(fx["beta"], prop, mol, ys_cond, batch["trg"][:, 1:].reshape(-1),
                                            mu, lv, False, 1)
        h = fx["hist"][step - 1]
        assert abs(float(loss) - h["loss"]) <= 5e-5 * abs(h["loss"])
        for p in params.values():
            p.grad = None
        loss.backward()
        with torch.no_grad():
            for k, (m, v) in state.items():
                O.adam_step(params[k], params[k].grad, m, v, step, lr)
        lr = O.noam_lr(step, cfg.d_model, fx["warmup"])
        assert abs(lr - h["lr"]) < 1e-12
    for k, full in fx["final_full"].items():
        # k_linear.bias has a mathematically zero gradient (softmax is invariant to a per-query
        # constant), so its fp32 gradient is rounding noise that Adam normalises to +-lr: chaotic.
        if k.endswith("pe.pe") or k.endswith("k_linear.bias"):
            continue
        assert rel_err(params[k], full) < 1e-4, k
