"""Operator-level parity (SIMT GEMM, Norm, attention, masks, loss) against plain torch fp32."""
import math

import pytest
import torch

import gct_plus_b200._lib as L
from gpu_common import DEV, gemm
from helpers import O, load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True)])
def test_simt_gemm_fp32(a_mn, b_mn):
    M, N, K = 200, 136, 100
    A = torch.randn(M, K, device=DEV)
    B = torch.randn(N, K, device=DEV)
    ref = A.double() @ B.double().t()
    out, _, _ = gemm(A.t().contiguous() if a_mn else A, B.t().contiguous() if b_mn else B, M, N, K, a_mn=a_mn, b_mn=b_mn,
                     dtype="fp32")
    assert float((out - ref).abs().max() / ref.abs().max()) < 1e-5


def test_norm_fwd_bwd_matches_reference_formula():
    m = load_golden("misc")
    x = m["norm_x"].to(DEV).reshape(-1, 512).contiguous()
    al, bi = m["norm_alpha"].to(DEV), m["norm_bias"].to(DEV)
    y = torch.empty_like(x)
    L.check(L.lib().gct_norm_fwd(L.ptr(x), L.ptr(al), L.ptr(bi), L.ptr(y), None, x.size(0), 512, 0, L.stream_ptr()))
    assert torch.allclose(y.cpu().view(4, 7, 512), m["norm_y"], atol=3e-6)
    # backward against autograd of the oracle formula
    xr = x.clone().requires_grad_(True)
    ar, br = al.clone().requires_grad_(True), bi.clone().requires_grad_(True)
    dy = torch.randn_like(x)
    add = torch.randn_like(x)
    O.norm(xr, ar, br).backward(dy)
    dx, da, db = torch.empty_like(x), torch.zeros(512, device=DEV), torch.zeros(512, device=DEV)
    L.check(L.lib().gct_norm_bwd(L.ptr(x), L.ptr(al), L.ptr(dy), L.ptr(add), L.ptr(dx), L.ptr(da), L.ptr(db), x.size(0), 512,
                                 L.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.allclose(dx, xr.grad + add, atol=2e-5, rtol=1e-4)
    assert torch.allclose(da, ar.grad, atol=1e-4, rtol=1e-4) and torch.allclose(db, br.grad, atol=1e-4, rtol=1e-4)


def _attn_ref(q, k, v, mask):
    B, Lq, d = q.shape
    H = d // 64
    qh = q.view(B, Lq, H, 64).transpose(1, 2)
    kh = k.view(B, -1, H, 64).transpose(1, 2)
    vh = v.view(B, -1, H, 64).transpose(1, 2)
    o, p = O.attention(qh, kh, vh, 64, mask)
    return o.transpose(1, 2).reshape(B, Lq, d), p


@pytest.mark.parametrize("dtype,tol", [("fp32", 2e-5), ("bf16", 2e-2)])
@pytest.mark.parametrize("Lq,Lk,dense", [(17, 17, True), (33, 45, False), (104, 104, True), (5, 130, False)])
def test_attention_fwd_bwd(dtype, tol, Lq, Lk, dense):
    B, H, d = 3, 2, 128
    tdt = torch.float32 if dtype == "fp32" else torch.bfloat16
    q = torch.randn(B, Lq, d, device=DEV).to(tdt)
    k = torch.randn(B, Lk, d, device=DEV).to(tdt)
    v = torch.randn(B, Lk, d, device=DEV).to(tdt)
    if dense:
        mask = torch.tril(torch.ones(Lq, Lk, device=DEV, dtype=torch.bool)).expand(B, Lq, Lk).clone()
        mask[1, :, 3] = False
        mb, mr = Lq * Lk, Lk
    else:
        lens = torch.tensor([Lk, max(1, Lk // 2), max(1, Lk - 3)], device=DEV)
        mask = (torch.arange(Lk, device=DEV)[None, :] < lens[:, None]).view(B, 1, Lk)
        mb, mr = Lk, 0
    m8 = mask.to(torch.uint8).contiguous()
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    ref, pref = _attn_ref(qf, kf, vf, mask)
    out = torch.empty(B, Lq, d, device=DEV, dtype=tdt)
    lse = torch.empty(B, H, Lq, device=DEV)
    probs = torch.empty(B, H, Lq, Lk, device=DEV)
    dt = 0 if dtype == "fp32" else 1
    lib = L.lib()
    L.check(lib.gct_attention_fwd(L.ptr(q), d, L.ptr(k), d, L.ptr(v), d, L.ptr(m8), mb, mr, L.ptr(out), d, L.ptr(lse), L.ptr(probs),
                                  B, H, Lq, Lk, dt, L.stream_ptr()))
    torch.cuda.synchronize()
    assert float((out.float() - ref).abs().max()) < tol * max(1.0, float(ref.abs().max()))
    assert float((probs - pref).abs().max()) < tol
    dO = torch.randn(B, Lq, d, device=DEV).to(tdt)
    ref.backward(dO.float())
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    L.check(lib.gct_attention_bwd(L.ptr(q), d, L.ptr(k), d, L.ptr(v), d, L.ptr(m8), mb, mr, L.ptr(lse), L.ptr(out), d, L.ptr(dO), d, L.ptr(dq), d,
                                  L.ptr(dk), d, L.ptr(dv), d, B, H, Lq, Lk, dt, L.stream_ptr()))
    torch.cuda.synchronize()
    for got, want in ((dq, qf.grad), (dk, kf.grad), (dv, vf.grad)):
        assert float((got.float() - want).abs().max()) < 2 * tol * max(1.0, float(want.abs().max()))


@pytest.mark.parametrize("Lq,Lk,dense", [(81, 81, False), (80, 80, True), (80, 81, False), (96, 33, False), (20, 96, False), (64, 64, True),
                                           (101, 101, False), (100, 100, True), (100, 101, False), (112, 97, False)])
@pytest.mark.parametrize("sm_budget", [0, 6])
def test_persistent_attention_matches_one_tile_per_cta(Lq, Lk, dense, sm_budget):
    """L <= 96 (forward) / L <= 112 (backward, one (Q, K) pair above 96) runs the persistent kernels (a CTA walks a range of
    (batch, head) tiles and prefetches the next tile's operands into dead shared-memory units); per tile they do the arithmetic of the one-tile-per-CTA kernels in the same order, so the outputs
    must be bit-identical -- with few SMs a CTA walks dozens of tiles (every buffer rotation and barrier phase is exercised)."""
    B, H, d = 67, 8, 512
    torch.manual_seed(Lq * 131 + Lk)
    q, dO = (torch.randn(B, Lq, d, device=DEV).bfloat16() for _ in range(2))
    k, v = (torch.randn(B, Lk, d, device=DEV).bfloat16() for _ in range(2))
    if dense:
        mask = torch.tril(torch.ones(Lq, Lk, device=DEV, dtype=torch.bool)).expand(B, Lq, Lk).clone()
        mask[:, :, 3] = torch.rand(B, device=DEV)[:, None] < 0.5
        mb, mr = Lq * Lk, Lk
    else:
        lens = torch.randint(1, Lk + 1, (B,), device=DEV)
        mask = (torch.arange(Lk, device=DEV)[None, :] < lens[:, None]).view(B, 1, Lk)
        mb, mr = Lk, 0
    m8 = mask.to(torch.uint8).contiguous()
    lib = L.lib()

    def run(persistent):
        lib.gct_set_attention_persistent(3 if persistent else 0)
        lib.gct_set_sm_budget(sm_budget if persistent else 0)
        out = torch.zeros(B, Lq, d, device=DEV, dtype=torch.bfloat16)
        lse = torch.zeros(B, H, Lq, device=DEV)
        probs = torch.zeros(B, H, Lq, Lk, device=DEV)
        dq, dk, dv = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
        try:
            L.check(lib.gct_attention_fwd(L.ptr(q), d, L.ptr(k), d, L.ptr(v), d, L.ptr(m8), mb, mr, L.ptr(out), d, L.ptr(lse), L.ptr(probs),
                                          B, H, Lq, Lk, 1, L.stream_ptr()))
            L.check(lib.gct_attention_bwd(L.ptr(q), d, L.ptr(k), d, L.ptr(v), d, L.ptr(m8), mb, mr, L.ptr(lse), L.ptr(out), d, L.ptr(dO), d,
                                          L.ptr(dq), d, L.ptr(dk), d, L.ptr(dv), d, B, H, Lq, Lk, 1, L.stream_ptr()))
            torch.cuda.synchronize()
        finally:
            lib.gct_set_attention_persistent(2)
            lib.gct_set_sm_budget(0)
        return out, lse, probs, dq, dk, dv

    got, want = run(1), run(0)
    for name, g, w in zip(("out", "lse", "probs", "dq", "dk", "dv"), got, want):
        assert torch.equal(g, w), f"{name}: max abs diff {float((g.float() - w.float()).abs().max())}"
    # and the pair against the oracle (fp32 autograd)
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    ref, _ = _attn_ref(qf, kf, vf, mask)
    ref.backward(dO.float())
    assert float((got[0].float() - ref).abs().max()) < 2e-2 * max(1.0, float(ref.abs().max()))
    for g, w in ((got[3], qf.grad), (got[4], kf.grad), (got[5], vf.grad)):
        assert float((g.float() - w).abs().max()) < 4e-2 * max(1.0, float(w.abs().max()))


@pytest.mark.parametrize("dtype,tol", [("fp32", 2e-5), ("bf16", 2e-2)])
@pytest.mark.parametrize("Lq,Lk", [(9, 21), (40, 81)])
def test_attention_row_without_a_visible_key(dtype, tol, Lq, Lk):
    """The reference fills masked scores with -1e9 (Model/sublayers.py:33-35): a query whose keys are ALL masked gets the uniform
    distribution over the Lk keys, its value gradient is dO / Lk and no gradient reaches q / k.  Never happens with valid batches
    (the first key is <sos> or a condition row); the kernels follow the reference anyway."""
    B, H, d = 3, 2, 128
    tdt = torch.float32 if dtype == "fp32" else torch.bfloat16
    torch.manual_seed(Lq + Lk)
    q, dO = (torch.randn(B, Lq, d, device=DEV).to(tdt) for _ in range(2))
    k, v = (torch.randn(B, Lk, d, device=DEV).to(tdt) for _ in range(2))
    lens = torch.tensor([Lk, 0, max(1, Lk - 3)], device=DEV)
    mask = (torch.arange(Lk, device=DEV)[None, :] < lens[:, None]).view(B, 1, Lk)
    m8 = mask.to(torch.uint8).contiguous()
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    ref, pref = _attn_ref(qf, kf, vf, mask)
    ref.backward(dO.float())
    out = torch.empty(B, Lq, d, device=DEV, dtype=tdt)
    lse = torch.empty(B, H, Lq, device=DEV)
    probs = torch.empty(B, H, Lq, Lk, device=DEV)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    dt = 0 if dtype == "fp32" else 1
    lib = L.lib()
    L.check(lib.gct_attention_fwd(L.ptr(q), d, L.ptr(k), d, L.ptr(v), d, L.ptr(m8), Lk, 0, L.ptr(out), d, L.ptr(lse), L.ptr(probs), B, H, Lq, Lk,
                                  dt, L.stream_ptr()))
    L.check(lib.gct_attention_bwd(L.ptr(q), d, L.ptr(k), d, L.ptr(v), d, L.ptr(m8), Lk, 0, L.ptr(lse), L.ptr(out), d, L.ptr(dO), d, L.ptr(dq), d,
                                  L.ptr(dk), d, L.ptr(dv), d, B, H, Lq, Lk, dt, L.stream_ptr()))
    torch.cuda.synchronize()
    assert float((probs[1] - 1.0 / Lk).abs().max()) < 1e-6
    assert float((out.float() - ref.detach()).abs().max()) < tol * max(1.0, float(ref.abs().max()))
    assert float((probs - pref.detach()).abs().max()) < tol
    assert float(dq[1].float().abs().max()) == 0.0 and float(dk[1].float().abs().max()) == 0.0
    for got, want in ((dq, qf.grad), (dk, kf.grad), (dv, vf.grad)):
        assert float((got.float() - want).abs().max()) < 2 * tol * max(1.0, float(want.abs().max()))


def test_masks_match_golden():
    from gct_plus_b200.Model.modules import get_src_mask, get_trg_mask
    m = load_golden("misc")
    t = m["mask_in"].to(DEV)
    conds = torch.zeros(2, 3, device=DEV)
    assert torch.equal(get_src_mask(t, 1).cpu(), m["src_mask"])
    assert torch.equal(get_src_mask(t, 1, conds).cpu(), m["src_mask_c"])
    assert torch.equal(get_trg_mask(t, 1, False).cpu(), m["trg_mask"] != 0)
    assert torch.equal(get_trg_mask(t, 1, True, conds).cpu(), m["trg_mask_c2d"] != 0)


def test_loss_function_and_grads():
    from gct_plus_b200.Train.trainer1 import loss_function
    torch.manual_seed(0)
    B, T, V, Se, lat = 5, 9, 27, 11, 32
    logits = torch.randn(B, T, V, device=DEV, requires_grad=True)
    mu = torch.randn(B, Se, lat, device=DEV, requires_grad=True)
    lv = (0.3 * torch.randn(B, Se, lat, device=DEV)).requires_grad_(True)
    ys = torch.randint(0, V, (B * T,), device=DEV)
    ys[::4] = 1
    loss, rce, _, kld = loss_function(0.37, None, logits, None, ys, mu, lv, False, 1)
    loss.backward()
    l2, m2, v2 = (t.detach().clone().requires_grad_(True) for t in (logits, mu, lv))
    loss_o, rce_o, _, kld_o = O.loss_function(0.37, None, l2, None, ys, m2, v2, False, 1)
    loss_o.backward()
    assert abs(float(loss) - float(loss_o)) < 1e-4 * abs(float(loss_o))
    assert abs(float(rce) - float(rce_o)) < 1e-4 * abs(float(rce_o)) and abs(float(kld) - float(kld_o)) < 1e-4 * abs(float(kld_o))
    assert torch.allclose(logits.grad, l2.grad, atol=1e-5) and torch.allclose(mu.grad, m2.grad, atol=1e-5)
    assert torch.allclose(lv.grad, v2.grad, atol=1e-5)
