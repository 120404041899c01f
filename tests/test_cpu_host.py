"""CPU-side checks (no GPU): the C-ABI library loads and exports every declared symbol, the module tree
reproduces the reference's state_dict / init, host-side helpers follow the reference."""
import os
import re

import numpy as np
import pytest
import torch

from helpers import load_golden, O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gct_plus_b200", "libgct_b200.so")


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(LIB):
        import __graft_entry__ as g
        g.build()
    import gct_plus_b200._lib as L
    return L


def test_library_exports_every_header_symbol(built):
    L = built
    lib = L.lib()
    hdr = open(os.path.join(ROOT, "include", "gct_b200.h")).read()
    declared = set(re.findall(r"^(?:int|int64_t|size_t|double|const char\*)\s+(gct_[a-z0-9_]+)\s*\(", hdr, flags=re.M))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/gct_b200.h but not exported"
    assert set(L.EXPORTED_SYMBOLS) <= declared
    assert lib.gct_sm() == 100 and lib.gct_version() >= 100
    assert lib.gct_num_slots(6) == 22 + 6 * 32
    assert abs(lib.gct_noam_lr(10, 512, 8000) - O.noam_lr(10, 512, 8000)) < 1e-15


def test_cpu_tensors_are_rejected_loudly(built):
    from gct_plus_b200.Model.modules import get_src_mask
    with pytest.raises(built.GctError):
        get_src_mask(torch.zeros(2, 3, dtype=torch.long), 1)


ARCH_FULL = dict(N=6, d_model=512, dff=2048, h=8, latent_dim=128)
ARCH_SMALL = dict(N=2, d_model=128, dff=256, h=2, latent_dim=32)
CASES = {"vaetf_full": ("vaetf", ARCH_FULL, 0, False, False), "pvaetf_full": ("pvaetf", ARCH_FULL, 3, False, True),
         "scavaetf_small": ("scavaetf", ARCH_SMALL, 0, False, False), "pscavaetf_small": ("pscavaetf", ARCH_SMALL, 3, False, True),
         "pvaetf_c2d_small": ("pvaetf", ARCH_SMALL, 3, True, False), "pvaetf_plain_small": ("pvaetf", ARCH_SMALL, 3, False, False)}


@pytest.mark.parametrize("name", list(CASES))
def test_init_matches_reference(name):
    """Same seed -> same state_dict keys, order, shapes and values as the reference's constructors."""
    from gct_plus_b200.Model import Cvaetf, Vaetf
    init = load_golden("init_parity")
    mt, arch, nc, c2d, c2l = CASES[name]
    torch.manual_seed(0)
    m = (Vaetf if mt == "vaetf" else Cvaetf)(32, 32, dropout=0.1, nconds=nc, use_cond2dec=c2d, use_cond2lat=c2l, **arch)
    sd, gold = m.state_dict(), init[name]
    assert list(sd.keys()) == list(gold.keys())
    for k, v in sd.items():
        g = gold[k]
        assert v.numel() == g["numel"], k
        assert torch.equal(v.flatten()[:8].float(), g["head"]), k
        assert abs(float(v.double().sum()) - g["sum"]) <= 1e-6 * max(1.0, abs(g["sum"])), k
    assert sum(p.numel() for p in m.parameters()) == init[name + "/nparams"]
    # parameters are views of one flat buffer in kernel layout; q|k|v rows are adjacent
    a = m.encoder.layers[0].attn
    assert a.k_linear.weight.data_ptr() == a.q_linear.weight.data_ptr() + a.q_linear.weight.numel() * 4
    assert a.v_linear.weight.data_ptr() == a.k_linear.weight.data_ptr() + a.k_linear.weight.numel() * 4


def test_state_dict_roundtrip_and_module_prefix(tmp_path):
    from gct_plus_b200.Model import Cvaetf
    from gct_plus_b200.Model.build_model import load_state
    torch.manual_seed(1)
    a = Cvaetf(32, 32, nconds=3, use_cond2lat=True, dropout=0.1, **ARCH_SMALL)
    torch.manual_seed(2)
    b = Cvaetf(32, 32, nconds=3, use_cond2lat=True, dropout=0.1, **ARCH_SMALL)
    path = str(tmp_path / "model_1.pt")
    torch.save({"model_state_dict": {"module." + k: v for k, v in a.state_dict().items()}}, path)
    load_state(b, path, 0)
    for (k, x), (_, y) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(x, y), k
    # still one flat buffer after load_state_dict (copy_ keeps the aliasing)
    assert b.encoder.layers[0].attn.q_linear.weight.data_ptr() == b._flat.data_ptr() + 4 * int(b._offsets[22 + 2])


def test_toklen_sampler_and_pe_table_are_bit_identical():
    from gct_plus_b200.Inference.toklen_sampling import tokenlen_gen_from_data_distribution
    from gct_plus_b200.Model.modules import positional_table
    m = load_golden("misc")
    np.random.seed(11)
    d = m["toklen_data"]
    out = tokenlen_gen_from_data_distribution(data=d, size=64, nBins=int(d.max() - d.min()))
    assert np.array_equal(out, m["toklen_out"])
    pe = positional_table(512)[0]
    assert torch.equal(pe[[0, 1, 2, 57, 199]], m["pe512_rows"])
    assert torch.equal(positional_table(128)[0][[0, 1, 99]], m["pe128_rows"])


def test_kl_annealer():
    from gct_plus_b200.Train.trainer1 import KLAnnealer
    m = load_golden("misc")
    assert [KLAnnealer(e, 0.02, 0.02, 1) for e in range(1, 6)] == m["kla"]


def test_batch_detokeniser_matches_id_to_smi(built):
    """ids_to_smiles (host-side gct_detokenize) == the reference's per-row id_to_smi (Inference/sampling_tool.py:54-61),
    including multi-character / multi-byte tokens, rows without <eos>, and the newline-in-token fallback."""
    import numpy as np
    from gct_plus_b200.Inference.sampling_tool import Sampling

    class _V:
        pass

    for vocab in (["<unk>", "<pad>", "<sos>", "<eos>", "<sep>", "C", "c", "Cl", "Br", "[nH]", "(", ")", "1", "=", "é"],
                  ["<unk>", "<pad>", "<sos>", "<eos>", "x\ny", "C"]):
        s = Sampling.__new__(Sampling)
        s.TRG = _V(); s.TRG.vocab = _V(); s.TRG.vocab.itos = vocab
        s._itos = np.array(vocab, dtype=object)
        s.sos_id, s.eos_id = 2, 3
        rng = np.random.RandomState(0)
        outs = rng.randint(0, len(vocab), size=(257, 40))
        outs[:, 0] = 2
        outs[5] = np.where(outs[5] == 3, 0, outs[5])          # a row that never emits <eos>
        assert s.ids_to_smiles(outs) == [s.id_to_smi(r) for r in outs]
    assert s.ids_to_smiles(np.zeros((0, 7), dtype=np.int64)) == []


def test_dropout_hash_statistics():
    """NumPy mirror of common.cuh's dropout decision (x = seed + pair*0x9e3779b9; x ^= x >> 15; x *= 0x2c1b3c6d; the two
    16-bit halves against thresh >> 16): keep rate, cross-half, lag-1 and row-stride correlations stay inside 4.5 sigma."""
    import numpy as np
    n, p = 1 << 21, 0.1
    t16 = int(min(4294967295.0, p * 4294967296.0)) >> 16
    for seed in (12345, 0x9E3779B9, 7):
        x = (seed + np.arange(n, dtype=np.uint64) * 0x9E3779B9) & 0xFFFFFFFF
        x ^= x >> 15
        x = (x * 0x2C1B3C6D) & 0xFFFFFFFF
        lo, hi = (x & 0xFFFF) < t16, (x >> 16) < t16
        sig = np.sqrt(p * (1 - p) / n)
        for d in (lo, hi):
            assert abs(d.mean() - t16 / 65536) < 4.5 * sig
        c = lambda a, b: abs(np.corrcoef(a.astype(float), b.astype(float))[0, 1]) * np.sqrt(n)   # noqa: E731
        assert c(lo, hi) < 4.5 and c(lo[:-1], lo[1:]) < 4.5 and c(hi[:-1], hi[1:]) < 4.5
        assert c(lo[:-256], lo[256:]) < 4.5 and c(hi[:-1024], hi[1024:]) < 4.5 and c(lo[:-1], hi[1:]) < 4.5


def test_flat_storage_survives_deepcopy_pickle_and_data_reassignment(tmp_path):
    """The kernels read one flat buffer; parameters must keep aliasing it after copy.deepcopy, torch.save(model) /
    torch.load and ``p.data = other`` (sync_weights re-flattens), and gradient views are keyed by parameter index."""
    import copy
    from gct_plus_b200.Model import Cvaetf
    torch.manual_seed(1)
    a = Cvaetf(32, 32, nconds=3, use_cond2lat=True, dropout=0.1, **ARCH_SMALL)
    assert a._aliased()
    b = copy.deepcopy(a)
    assert b._flat.data_ptr() != a._flat.data_ptr() and b._aliased()
    assert torch.equal(b._flat, a._flat)
    assert len(b._grad_views) == len(list(b.parameters()))
    with torch.no_grad():
        b.out.weight.mul_(2.0)
    off = b._grad_views[[n for n, _ in b.named_parameters()].index("out.weight")][0]
    assert torch.equal(b._flat[off:off + b.out.weight.numel()].view_as(b.out.weight), b.out.weight)   # write lands in _flat
    assert not torch.equal(a.out.weight, b.out.weight)                                               # ... and only in the copy's
    path = str(tmp_path / "whole_model.pt")
    torch.save(a, path)
    c = torch.load(path, weights_only=False)
    assert c._aliased() and torch.equal(c._flat, a._flat)
    # out-of-band re-pointing of a parameter: detected, repaired, new values reach the flat buffer
    a.out.bias.data = torch.full_like(a.out.bias, 3.0)
    assert not a._aliased()
    a.sync_weights()
    assert a._aliased()
    offb = a._grad_views[[n for n, _ in a.named_parameters()].index("out.bias")][0]
    assert float(a._flat[offb]) == 3.0
    # a load_state_dict into the deep copy changes what the kernels would read
    b.load_state_dict(a.state_dict())
    assert torch.equal(b._flat, a._flat)


def test_fused_trainer_property_head_configuration():
    """use_cond2dec adds prop_fc + an MSE term and shifts the target rows (trainer1.py:24-26): FusedTrainer takes it from the
    model and refuses an inconsistent request instead of training on a wrong loss."""
    from gct_plus_b200._lib import GctError
    from gct_plus_b200.Model import Cvaetf
    from gct_plus_b200.Train.trainer1 import FusedTrainer
    m = Cvaetf(32, 32, nconds=3, use_cond2dec=True, dropout=0.1, **ARCH_SMALL)
    assert FusedTrainer(m, "pvaetf").use_cond2dec
    with pytest.raises(GctError):
        FusedTrainer(m, "scavaetf")                      # property rows without a property-conditioned model type
    m2 = Cvaetf(32, 32, nconds=3, use_cond2lat=True, dropout=0.1, **ARCH_SMALL)
    assert not FusedTrainer(m2, "pvaetf").use_cond2dec
    with pytest.raises(GctError):
        FusedTrainer(m2, "pvaetf", use_cond2dec=True)    # the model has no property head


def test_host_toklen_draw_continues_numpys_global_stream(built):
    """gct_toklen_draw (C loop in the library) == the reference's per-draw Python loop bit for bit, including a pending
    cached Gaussian on entry, and leaves np.random in the identical state (the next draws agree)."""
    from gct_plus_b200.Inference import toklen_sampling as T
    data = load_golden("misc")["toklen_data"]
    nb = int(data.max() - data.min())
    counts, edges = np.histogram(data, bins=nb)
    width = np.diff(edges)[0]
    centres = edges[:-1] + 0.5 * width
    cdf = np.zeros_like(edges)
    cdf[1:] = np.cumsum(counts / np.sum(counts))
    for seed, size, pending in ((11, 64, False), (3, 5001, True), (5, 1, True), (9, 0, False)):
        res = []
        for fn in (lambda: T.tokenlen_gen_from_data_distribution(data, size, nb), lambda: T._python_loop(cdf, centres, width, size)):
            np.random.seed(seed)
            if pending:
                np.random.normal()              # leaves the second value of the polar pair cached
            out = fn()
            res.append((out, np.random.uniform(), np.random.normal(), np.random.randint(1 << 30)))
        assert np.array_equal(res[0][0], res[1][0]) and res[0][1:] == res[1][1:], (seed, size)


def test_gradient_buckets_cover_the_flat_buffer_once_and_follow_the_backward_order():
    """Bucket table of the overlapped DP exchange (gct_backward_dp): every element of the flat gradient buffer belongs to
    exactly one bucket, a layer's parameters sit in that layer's bucket, and stages run decoder N-1..0, encoder N-1..0, rest."""
    import gct_plus_b200._lib as L
    from gct_plus_b200.Model import Cvaetf, Vaetf
    from gct_plus_b200.Train.dp import gradient_buckets
    for m in (Cvaetf(32, 32, nconds=3, use_cond2lat=True, dropout=0.1, **ARCH_SMALL), Vaetf(32, 32, nconds=0, dropout=0.1, **ARCH_SMALL)):
        N, total = len(m.encoder.layers), m._flat.numel()
        bk = gradient_buckets(m._offsets, N, total, L.NUM_GLOBAL_SLOTS, L.ENC_LAYER_SLOTS, L.DEC_LAYER_SLOTS)
        cover = np.zeros(total, dtype=np.int32)
        stage_of = np.full(total, -1)
        for stage, off, cnt in bk:
            cover[off:off + cnt] += 1
            stage_of[off:off + cnt] = stage
        assert (cover == 1).all()
        for (name, _), (off, n, _) in zip(m.named_parameters(), m._grad_views):
            st = set(stage_of[off:off + n].tolist())
            assert len(st) == 1, name
            st = st.pop()
            parts = name.split(".")
            if parts[1] == "layers":
                l = int(parts[2])
                assert st == (N - 1 - l if parts[0] == "decoder" else 2 * N - 1 - l), name
            else:
                assert st == 2 * N, name


def test_mt19937_fill_continues_torchs_cpu_generator(built):
    """gct_mt19937_fill on a copy of torch.get_rng_state(): the raw outputs give exactly torch.rand's uniforms and the engine
    ends in exactly torch's state, from any position inside a 624-word block and for any count."""
    lib = built.lib()
    for pre in (0, 1, 623, 624, 777):
        torch.manual_seed(5)
        if pre:
            torch.rand(pre)
        for n in (1, 15, 16, 623, 624, 625, 5000, 1248 + 3):
            keep = torch.get_rng_state().clone()
            sn = keep.clone().numpy()
            raw = np.empty(n, dtype=np.uint32)
            built.check(lib.gct_mt19937_fill(sn[24:24 + 4992].view(np.uint64).ctypes.data, sn[8:12].view(np.int32).ctypes.data,
                                             sn[16:24].view(np.uint64).ctypes.data, raw.ctypes.data, n))
            u = torch.rand(n).numpy()
            assert np.array_equal((raw & 0xFFFFFF).astype(np.float32) * np.float32(2.0 ** -24), u), (pre, n)
            assert np.array_equal(torch.get_rng_state().numpy()[:5016], sn[:5016]), (pre, n)
            torch.set_rng_state(keep)
