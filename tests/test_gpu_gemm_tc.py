"""tcgen05 GEMM (gemm_tc.cuh) against torch fp32 matmul on the same bf16-rounded operands.
Tolerance 2e-3 relative to the output scale: only the fp32 accumulation order differs."""
import pytest
import torch

from gpu_common import DEV, gemm

pytestmark = pytest.mark.gpu


def _ops(M, N, K, a_mn, b_mn, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    A = torch.randn(M, K, generator=g).to(DEV).bfloat16()
    B = torch.randn(N, K, generator=g).to(DEV).bfloat16()
    ref = A.float() @ B.float().t()
    As = A.t().contiguous() if a_mn else A
    Bs = B.t().contiguous() if b_mn else B
    return As, Bs, ref


def _close(out, ref, tol=2e-3):
    err = float((out.float() - ref).abs().max() / ref.abs().max())
    assert err < tol, err


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("M,N,K,bn", [(128, 128, 64, 0), (256, 128, 128, 0), (512, 1536, 512, 0), (296, 96, 200, 64),
                                      (512, 512, 2048, 32), (1024, 256, 512, 256), (384, 64, 320, 64), (512, 512, 512, 16), (128, 48, 64, 16)])
def test_majorness(a_mn, b_mn, M, N, K, bn):
    if b_mn and bn in (16, 32):
        bn = 64
    if a_mn and bn == 16:
        bn = 32
    As, Bs, ref = _ops(M, N, K, a_mn, b_mn)
    out, _, _ = gemm(As, Bs, M, N, K, a_mn=a_mn, b_mn=b_mn, bn=bn)
    _close(out, ref)


def _gelu_grad(x):
    return 0.5 * (1 + torch.erf(x * 0.7071067811865476)) + x * torch.exp(-0.5 * x * x) * 0.3989422804014327


@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
def test_gelu_grad_and_mul_aux_small(dtype):
    """Non-persistent / SIMT tiers of the two epilogue modes the FFN forward / backward use."""
    M, N, K = 256, 256, 128
    As, Bs, ref = _ops(M, N, K, False, False, seed=11)
    bias = torch.randn(N, device=DEV)
    tol = 1e-2 if dtype == "bf16" else 1e-5
    A_, B_ = (As, Bs) if dtype == "bf16" else (As.float(), Bs.float())
    _, outT, aux = gemm(A_, B_, M, N, K, bias=bias, flags=1 | 128, want_T=True, dtype=dtype)
    _close(outT, torch.nn.functional.gelu(ref + bias), tol)
    _close(aux, _gelu_grad(ref + bias), tol)
    aux_in = torch.randn(M, N, device=DEV).to(outT.dtype)
    _, out2, _ = gemm(A_, B_, M, N, K, flags=256, want_T=True, aux_in=aux_in, dtype=dtype)
    _close(out2, ref * aux_in.float(), tol)
    _, out3, _ = gemm(A_, B_, M, N, K, flags=2, want_T=True, aux_in=aux_in, dtype=dtype)
    _close(out3, ref * _gelu_grad(aux_in.float()), tol)


def test_small_vocab_tile():
    As, Bs, ref = _ops(512, 28, 512, False, False, seed=3)      # N = 28 < BN = 32, ldc = 28
    out, _, _ = gemm(As, Bs, 512, 28, 512)
    _close(out, ref)


def test_epilogue_bias_gelu_residual():
    M, N, K = 256, 256, 128
    As, Bs, ref = _ops(M, N, K, False, False, seed=1)
    bias = torch.randn(N, device=DEV)
    res = torch.randn(M, N, device=DEV)
    out, _, _ = gemm(As, Bs, M, N, K, bias=bias, res=res)
    _close(out, ref + bias + res)
    out2, outT, aux = gemm(As, Bs, M, N, K, bias=bias, flags=1, want_T=True)
    pre = ref + bias
    _close(aux, pre, 1e-2)
    _close(outT, torch.nn.functional.gelu(pre), 1e-2)


def test_split_k_accumulate():
    M, N, K = 256, 512, 4096
    As, Bs, ref = _ops(M, N, K, True, True, seed=2)
    acc = torch.ones((M, N), device=DEV)
    gemm(As, Bs, M, N, K, a_mn=True, b_mn=True, flags=4, split_k=8, out32=acc)
    _close(acc, ref + 1.0)


@pytest.mark.parametrize("N,K", [(512, 512), (1536, 512), (256, 512), (2048, 128), (384, 192)])
@pytest.mark.parametrize("mode", ["bias_T", "bias_res_F32", "gelu", "gelu_grad", "mul_aux", "dgelu", "plain_F32", "plain_T", "bias_F32", "res_F32"])
def test_persistent_kernel_epilogue_modes(N, K, mode):
    """More tiles than SMs -> persistent kernel with the specialised, smem-staged epilogue; every mode the model uses."""
    M = 19 * 1024 + 128          # 153 full row tiles (+ a partial-tile run below)
    for Mx in (M, M - 40):
        As, Bs, ref = _ops(Mx, N, K, False, False, seed=7)
        bias = torch.randn(N, device=DEV)
        res = torch.randn(Mx, N, device=DEV)
        if mode == "bias_T":
            _, outT, _ = gemm(As, Bs, Mx, N, K, bias=bias, want_T=True)
            _close(outT, ref + bias, 1e-2)
        elif mode == "bias_res_F32":
            out, _, _ = gemm(As, Bs, Mx, N, K, bias=bias, res=res)
            _close(out, ref + bias + res)
        elif mode == "gelu":
            _, outT, aux = gemm(As, Bs, Mx, N, K, bias=bias, flags=1, want_T=True)
            _close(aux, ref + bias, 1e-2)
            _close(outT, torch.nn.functional.gelu(ref + bias), 1e-2)
        elif mode == "gelu_grad":       # EPI_GELU | EPI_GELU_GRAD: aux = gelu'(pre) (dropout off -> keep = 1)
            _, outT, aux = gemm(As, Bs, Mx, N, K, bias=bias, flags=1 | 128, want_T=True)
            _close(outT, torch.nn.functional.gelu(ref + bias), 1e-2)
            _close(aux, _gelu_grad(ref + bias), 1e-2)
        elif mode == "mul_aux":         # EPI_MUL_AUX: out = acc * aux_in
            aux_in = torch.randn(Mx, N, device=DEV).bfloat16()
            _, outT, _ = gemm(As, Bs, Mx, N, K, flags=256, want_T=True, aux_in=aux_in)
            _close(outT, ref * aux_in.float(), 1e-2)
        elif mode == "dgelu":           # EPI_DGELU: out = acc * gelu'(aux_in)   (pre-activation form of the FFN backward)
            aux_in = torch.randn(Mx, N, device=DEV).bfloat16()
            _, outT, _ = gemm(As, Bs, Mx, N, K, flags=2, want_T=True, aux_in=aux_in)
            _close(outT, ref * _gelu_grad(aux_in.float()), 1e-2)
        elif mode == "plain_F32":
            out, _, _ = gemm(As, Bs, Mx, N, K)
            _close(out, ref)
        elif mode == "plain_T":
            _, outT, _ = gemm(As, Bs, Mx, N, K, want_T=True)
            _close(outT, ref, 1e-2)
        elif mode == "bias_F32":
            out, _, _ = gemm(As, Bs, Mx, N, K, bias=bias)
            _close(out, ref + bias)
        else:
            out, _, _ = gemm(As, Bs, Mx, N, K, res=res)
            _close(out, ref + res)


@pytest.fixture
def _pair_switch():
    import gct_plus_b200._lib as L
    yield L.lib().gct_set_cta_pair_gemm
    L.lib().gct_set_cta_pair_gemm(2)


@pytest.mark.parametrize("pair", [2, 0])
@pytest.mark.parametrize("M,N,K", [(19584, 512, 512), (19584 - 40, 1536, 512), (41472, 2048, 512), (30000, 512, 2048)])
def test_cta_pair_kmajor_at_training_shapes(_pair_switch, pair, M, N, K):
    """CTA-pair (tcgen05.mma.cta_group::2) persistent GEMM at the cfg 3 / cfg 4 row counts against torch fp32 on the same
    bf16 operands; pair=0 runs the single-CTA persistent kernel on the same inputs."""
    _pair_switch(pair)
    As, Bs, ref = _ops(M, N, K, False, False, seed=5)
    bias = torch.randn(N, device=DEV)
    res = torch.randn(M, N, device=DEV)
    _, outT, _ = gemm(As, Bs, M, N, K, bias=bias, want_T=True)
    _close(outT, ref + bias, 1e-2)
    out, _, _ = gemm(As, Bs, M, N, K, bias=bias, res=res)
    _close(out, ref + bias + res)
    _, g, aux = gemm(As, Bs, M, N, K, bias=bias, flags=1 | 128, want_T=True)
    _close(g, torch.nn.functional.gelu(ref + bias), 1e-2)
    _close(aux, _gelu_grad(ref + bias), 1e-2)


@pytest.mark.parametrize("pair", [2, 0])
@pytest.mark.parametrize("M,N,K,a_mn,b_mn,split", [
    (41472, 512, 1536, False, True, 1),      # dgrad of the QKV projection at cfg 3 (B = 512, Se = 81): gemm_tc_persist_kernel<256,0,1,..>
    (41472, 512, 2048, False, True, 1),      # dgrad of FFN linear_1
    (41472, 2048, 512, False, True, 1),      # dgrad of FFN linear_2 (multiply-by-aux epilogue covered separately)
    (1536, 512, 41472, True, True, 9),       # split-K wgrad of the QKV projection: gemm_tc_persist_kernel<256,1,1,..>
    (2048, 512, 41472, True, True, 5),       # split-K wgrad of FFN linear_1
    (512, 2048, 41472, True, True, 5),       # split-K wgrad of FFN linear_2
    (512, 512, 40448, True, True, 37),       # split-K wgrad of an out-projection (B*T = 40448 decoder rows)
    (512, 512, 51712, True, True, 37),       # cfg 4 encoder rows (B = 512, Se = 101)
])
def test_cta_pair_dgrad_and_split_k_wgrad(_pair_switch, pair, M, N, K, a_mn, b_mn, split):
    """The operand layouts of the backward pass (activation gradients with an MN-major weight, weight gradients with both
    operands MN-major and split-K accumulation into fp32) at the benched row counts, pair and single-CTA kernels."""
    _pair_switch(pair)
    As, Bs, ref = _ops(M, N, K, a_mn, b_mn, seed=9)
    if split == 1:
        out, _, _ = gemm(As, Bs, M, N, K, a_mn=a_mn, b_mn=b_mn)
        _close(out, ref)
        aux_in = torch.randn(M, N, device=DEV).bfloat16()
        _, outT, _ = gemm(As, Bs, M, N, K, a_mn=a_mn, b_mn=b_mn, flags=256, want_T=True, aux_in=aux_in)
        _close(outT, ref * aux_in.float(), 1e-2)
        res = torch.randn(M, N, device=DEV)
        out, _, _ = gemm(As, Bs, M, N, K, a_mn=a_mn, b_mn=b_mn, res=res)
        _close(out, ref + res)
    else:
        out = torch.ones(M, N, device=DEV)
        gemm(As, Bs, M, N, K, a_mn=a_mn, b_mn=b_mn, flags=4, split_k=split, out32=out)
        _close(out, ref + 1.0)


@pytest.mark.parametrize("M,K", [(300, 512), (128 * 150 + 17, 512), (4096, 1024), (1000, 1536), (2500, 2048)])
@pytest.mark.parametrize("variant", ["full", "inplace", "no_bias_res", "norm32"])
def test_residual_projection_fused_with_norm(M, K, variant):
    """gct_gemm_rownorm (128 x 512 tiles, whole rows per CTA, Norm in the epilogue) against torch: x = A W^T + b + res and
    y = alpha (x - mean) / (std_unbiased + eps) + beta (Model/modules.py:80-95), ragged last tile, in-place residual stream."""
    import gct_plus_b200._lib as L
    lib = L.lib()
    g = torch.Generator().manual_seed(M + K)
    A = torch.randn(M, K, generator=g).to(DEV).bfloat16()
    W = (torch.randn(512, K, generator=g) / K ** 0.5).to(DEV).bfloat16()
    bias = torch.randn(512, generator=g).to(DEV) if variant != "no_bias_res" else None
    res = (2 * torch.randn(M, 512, generator=g) + 0.5).to(DEV) if variant != "no_bias_res" else None
    alpha = (1 + 0.1 * torch.randn(512, generator=g)).to(DEV)
    beta = (0.1 * torch.randn(512, generator=g)).to(DEV)
    x_ref = A.float() @ W.float().t()
    if bias is not None:
        x_ref = x_ref + bias + res
    y_ref = alpha * (x_ref - x_ref.mean(-1, keepdim=True)) / (x_ref.std(-1, keepdim=True) + 1e-6) + beta
    out32 = res.clone() if variant == "inplace" else torch.full((M, 512), 7.0, device=DEV)
    res_in = out32 if variant == "inplace" else res
    normT = torch.zeros(M, 512, device=DEV, dtype=torch.bfloat16)
    norm32 = torch.zeros(M, 512, device=DEV) if variant == "norm32" else None
    L.check(lib.gct_gemm_rownorm(L.ptr(A), K, L.ptr(W), K, M, K, L.ptr(bias), L.ptr(res_in), L.ptr(out32), L.ptr(alpha), L.ptr(beta),
                                 L.ptr(normT), L.ptr(norm32), 1e-6, L.stream_ptr()), "gct_gemm_rownorm")
    torch.cuda.synchronize()
    _close(out32, x_ref, 2e-3)
    _close(normT, y_ref, 1e-2)
    if norm32 is not None:
        _close(norm32, y_ref, 2e-3)
