"""CPU oracle for the GCT-Plus Transformer-VAE hot path -- TEST INFRASTRUCTURE ONLY.

This file is a functional restatement (plain torch fp32 ops over a ``state_dict``)
of the arithmetic the reference performs on the path named by BASELINE.json.  It is
the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The product package
(``gct_plus_b200``) never imports anything under ``oracle/`` and has no CPU fallback.

Parity pinning: the reference ships no tests or golden vectors, so
the oracle is pinned against outputs of the reference's own modules executed in the
build container (``oracle/make_golden.py`` imports ``/root/reference`` unmodified, with
the three import shims of SURVEY.md section 8c, and writes ``tests/golden/*.pt``).
``tests/test_oracle_golden.py`` checks every function here against those fixtures.

Every function cites the reference file:line it follows (paths relative to the
reference root).  All tensors are batch-first.  ``sd`` is a reference-format
``state_dict`` (keys as produced by ``Model/vaetf.py`` / ``Model/cvaetf.py``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


@dataclass
class ModelCfg:
    """Hyper-parameters of one model flavour (Model/build_model.py:42-56)."""
    model_type: str = "vaetf"          # vaetf | pvaetf | scavaetf | pscavaetf
    src_vocab: int = 32
    trg_vocab: int = 32
    N: int = 6
    d_model: int = 512
    dff: int = 2048
    h: int = 8
    latent_dim: int = 128
    nconds: int = 0
    use_cond2dec: bool = False
    use_cond2lat: bool = False

    @property
    def is_vaetf(self) -> bool:         # Vaetf keeps mu/log_var heads in `sampler.`, Cvaetf in `encoder.`
        return self.model_type == "vaetf"


# ----------------------------------------------------------------------------------------------
# a1  Norm                                                           Model/modules.py:80-95
# ----------------------------------------------------------------------------------------------
def norm(x: Tensor, alpha: Tensor, bias: Tensor, eps: float = 1e-6) -> Tensor:
    """alpha*(x-mean)/(std_unbiased+eps)+bias -- eps is added to sigma, sigma uses N-1."""
    mean = x.mean(dim=-1, keepdim=True)
    xc = x - mean
    var = (xc * xc).sum(dim=-1, keepdim=True) / (x.size(-1) - 1)
    return alpha * xc / (var.sqrt() + eps) + bias


# ----------------------------------------------------------------------------------------------
# a3  PositionalEncoding table                                       Model/modules.py:116-144
# ----------------------------------------------------------------------------------------------
def positional_table(max_seq_len: int, d_model: int) -> Tensor:
    """pe[pos,i]=sin(pos/10000^(2i/d)), pe[pos,i+1]=cos(pos/10000^(2(i+1)/d)), i even.

    Computed in float64 then rounded to fp32 exactly like the reference's python-float loop
    (modules.py:123-130)."""
    pos = np.arange(max_seq_len, dtype=np.float64)[:, None]
    i = np.arange(0, d_model, 2, dtype=np.float64)[None, :]
    pe = np.zeros((max_seq_len, d_model), dtype=np.float64)
    pe[:, 0::2] = np.sin(pos / (10000.0 ** ((2.0 * i) / d_model)))
    pe[:, 1::2] = np.cos(pos / (10000.0 ** ((2.0 * (i + 1.0)) / d_model)))
    return torch.from_numpy(pe.astype(np.float32))


def add_positional(x: Tensor, pe: Tensor, d_model: int) -> Tensor:
    """x*sqrt(d)+pe[:L]; dropout omitted (eval).                  Model/modules.py:136-143"""
    return x * math.sqrt(d_model) + pe[: x.size(1)].to(x.device)


# ----------------------------------------------------------------------------------------------
# a4  masks                                                          Model/modules.py:10-66
# ----------------------------------------------------------------------------------------------
def src_mask(src: Tensor, pad_idx: int, nconds: int = 0) -> Tensor:
    """(src!=pad)[:,None,:], with `nconds` leading True columns.   Model/modules.py:33-44"""
    m = (src != pad_idx).unsqueeze(-2)
    if nconds > 0:
        ones = torch.ones(src.size(0), 1, nconds, dtype=torch.bool, device=src.device)
        m = torch.cat([ones, m], dim=2)
    return m


def nopeak_mask(trg_size: int, use_cond2dec: bool, cond_dim: int = 0) -> Tensor:
    """Boolean 'may attend' matrix (1,L,L).                         Model/modules.py:17-30

    Without cond2dec: lower-triangular.  With cond2dec the (nc+T)x(nc+T) block matrix is
      [ all-True (nc x nc)  | only column 0 True (nc x T) ]
      [ all-True (T  x nc)  | lower-triangular   (T  x T) ]"""
    tri = torch.tril(torch.ones(trg_size, trg_size, dtype=torch.bool))
    if use_cond2dec:
        ul = torch.ones(cond_dim, cond_dim, dtype=torch.bool)
        ur = torch.zeros(cond_dim, trg_size, dtype=torch.bool)
        ur[:, 0] = True
        ll = torch.ones(trg_size, cond_dim, dtype=torch.bool)
        tri = torch.cat([torch.cat([ul, ur], dim=1), torch.cat([ll, tri], dim=1)], dim=0)
    return tri.unsqueeze(0)


def trg_mask(trg: Tensor, pad_idx: int, use_cond2dec: bool = False, nconds: int = 0) -> Tensor:
    """Key-padding AND no-peak.                                     Model/modules.py:47-58

    The reference multiplies the bool no-peak mask by pad_idx (an int) and ANDs it with the
    bool padding mask, which is only a logical AND because <pad> == 1; the result there is an
    int64 {0,1} tensor, here a bool tensor with the same truth values."""
    m = (trg != pad_idx).unsqueeze(-2)
    if use_cond2dec:
        ones = torch.ones(trg.size(0), 1, nconds, dtype=torch.bool, device=trg.device)
        m = torch.cat([ones, m], dim=2)
    npk = nopeak_mask(trg.size(1), use_cond2dec, nconds).to(trg.device)
    return m & npk


# ----------------------------------------------------------------------------------------------
# a5-a7  attention / MultiHeadAttention / FeedForward               Model/sublayers.py:29-89
# ----------------------------------------------------------------------------------------------
def linear(x: Tensor, sd: Dict[str, Tensor], prefix: str) -> Tensor:
    return x @ sd[prefix + ".weight"].t() + sd[prefix + ".bias"]


def attention(q: Tensor, k: Tensor, v: Tensor, d_k: int, mask: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    """softmax(masked_fill(QK^T/sqrt(dk), mask==0, -1e9)) V.        Model/sublayers.py:29-41"""
    scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(d_k)
    if mask is not None:
        scores = scores.masked_fill(mask.unsqueeze(1) == 0, -1e9)
    probs = torch.softmax(scores, dim=-1)
    return torch.matmul(probs, v), probs


def mha(sd: Dict[str, Tensor], prefix: str, h: int, q_in: Tensor, kv_in: Tensor,
        mask: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    """Model/sublayers.py:44-74 (dropout omitted)."""
    bs, d_model = q_in.size(0), q_in.size(-1)
    d_k = d_model // h
    k = linear(kv_in, sd, prefix + ".k_linear").view(bs, -1, h, d_k).transpose(1, 2)
    q = linear(q_in, sd, prefix + ".q_linear").view(bs, -1, h, d_k).transpose(1, 2)
    v = linear(kv_in, sd, prefix + ".v_linear").view(bs, -1, h, d_k).transpose(1, 2)
    o, probs = attention(q, k, v, d_k, mask)
    concat = o.transpose(1, 2).contiguous().view(bs, -1, d_model)
    return linear(concat, sd, prefix + ".out"), probs


def feed_forward(sd: Dict[str, Tensor], prefix: str, x: Tensor) -> Tensor:
    """linear_2(gelu_erf(linear_1(x))).                             Model/sublayers.py:77-89"""
    return linear(F.gelu(linear(x, sd, prefix + ".linear_1")), sd, prefix + ".linear_2")


# ----------------------------------------------------------------------------------------------
# a9-a10  layers                                                     Model/layers.py:8-82
# ----------------------------------------------------------------------------------------------
def encoder_layer(sd, p: str, h: int, x: Tensor, mask: Tensor) -> Tuple[Tensor, Tensor]:
    """Residual is taken on the *normalised* tensor (layers.py:23,28,32,34)."""
    x = norm(x, sd[p + ".norm_1.alpha"], sd[p + ".norm_1.bias"])
    a, probs = mha(sd, p + ".attn", h, x, x, mask)
    x = x + a
    x = norm(x, sd[p + ".norm_2.alpha"], sd[p + ".norm_2.bias"])
    x = x + feed_forward(sd, p + ".ff", x)
    return x, probs


def decoder_layer(sd, p: str, h: int, x: Tensor, mem: Tensor, smask: Tensor, tmask: Tensor):
    """Residual is taken on the raw stream (layers.py:59,64,68,73,77-78)."""
    x2 = norm(x, sd[p + ".norm_1.alpha"], sd[p + ".norm_1.bias"])
    a1, p1 = mha(sd, p + ".attn_1", h, x2, x2, tmask)
    x = x + a1
    x2 = norm(x, sd[p + ".norm_2.alpha"], sd[p + ".norm_2.bias"])
    a2, p2 = mha(sd, p + ".attn_2", h, x2, mem, smask)
    x = x + a2
    x2 = norm(x, sd[p + ".norm_3.alpha"], sd[p + ".norm_3.bias"])
    x = x + feed_forward(sd, p + ".ff", x2)
    return x, p1, p2


# ----------------------------------------------------------------------------------------------
# a8, a11-a13  Encoder / Sampler / Decoder / model          Model/vaetf.py, Model/cvaetf.py
# ----------------------------------------------------------------------------------------------
def _pe(sd, which: str) -> Tensor:
    return sd[which + ".pe.pe"][0]


def encoder_trunk(sd, cfg: ModelCfg, src: Tensor, smask: Tensor, econds: Optional[Tensor],
                  want_attn: bool = False):
    """embed (|| cond tokens) -> PE -> N layers -> Norm.   vaetf.py:33-54 / cvaetf.py:35-53"""
    x = sd["encoder.embed_sentence.embed.weight"][src]
    if cfg.nconds > 0:
        c = linear(econds, sd, "encoder.embed_cond2enc").view(econds.size(0), cfg.nconds, -1)
        x = torch.cat([c, x], dim=1)
    x = add_positional(x, _pe(sd, "encoder"), cfg.d_model)
    attn = []
    for i in range(cfg.N):
        x, pr = encoder_layer(sd, f"encoder.layers.{i}", cfg.h, x, smask)
        attn.append(pr)
    x = norm(x, sd["encoder.norm.alpha"], sd["encoder.norm.bias"])
    return (x, attn) if want_attn else x


def latent_heads(sd, cfg: ModelCfg, x: Tensor, eps: Optional[Tensor]):
    """mu, log_var, z = mu + eps*exp(0.5*log_var).  sublayers.py:7-26 / cvaetf.py:55-69.

    `eps` is the N(0,1) draw the reference takes with randn_like (always variational,
    SURVEY.md 3.4); pass None for z = mu."""
    head = "sampler" if cfg.is_vaetf else "encoder"
    mu = linear(x, sd, head + ".fc_mu")
    log_var = linear(x, sd, head + ".fc_log_var")
    z = mu if eps is None else eps * torch.exp(0.5 * log_var) + mu
    return z, mu, log_var


def encode(sd, cfg: ModelCfg, src, smask, econds=None, eps=None):
    """Vaetf.encode / Cvaetf.encode.                    vaetf.py:145-148 / cvaetf.py:169-171"""
    x = encoder_trunk(sd, cfg, src, smask, econds)
    return latent_heads(sd, cfg, x, eps)


def decoder_trunk(sd, cfg: ModelCfg, trg, z, smask, tmask, dconds=None, want_attn: bool = False):
    """vaetf.py:79-114 / cvaetf.py:94-133."""
    x = sd["decoder.embed.embed.weight"][trg]
    mem = linear(z, sd, "decoder.fc_z")
    if cfg.use_cond2dec and cfg.nconds > 0:
        c = linear(dconds, sd, "decoder.embed_cond2dec").view(dconds.size(0), cfg.nconds, -1)
        x = torch.cat([c, x], dim=1)
    elif cfg.use_cond2lat and cfg.nconds > 0:
        c = linear(dconds, sd, "decoder.embed_cond2lat").view(dconds.size(0), cfg.nconds, -1)
        mem = torch.cat([c, mem], dim=1)
    x = add_positional(x, _pe(sd, "decoder"), cfg.d_model)
    if cfg.use_cond2lat and cfg.nconds > 0:
        ones = torch.ones(smask.size(0), 1, cfg.nconds, dtype=torch.bool, device=smask.device)
        smask = torch.cat([ones, smask], dim=2)
    a1s, a2s = [], []
    for i in range(cfg.N):
        x, p1, p2 = decoder_layer(sd, f"decoder.layers.{i}", cfg.h, x, mem, smask, tmask)
        a1s.append(p1)
        a2s.append(p2)
    x = norm(x, sd["decoder.norm.alpha"], sd["decoder.norm.bias"])
    return (x, a1s, a2s) if want_attn else x


def decode_logits(sd, cfg: ModelCfg, trg, z, smask, tmask, dconds=None) -> Tensor:
    """Vaetf.decode / Cvaetf.decode.                    vaetf.py:150-152 / cvaetf.py:173-175"""
    return linear(decoder_trunk(sd, cfg, trg, z, smask, tmask, dconds), sd, "out")


def forward(sd, cfg: ModelCfg, src, trg, smask, tmask, econds=None, dconds=None, eps=None):
    """(output_prop, output_mol, mu, log_var, z).       vaetf.py:154-182 / cvaetf.py:177-193"""
    z, mu, log_var = encode(sd, cfg, src, smask, econds, eps)
    out = decode_logits(sd, cfg, trg, z, smask, tmask, dconds)
    if cfg.use_cond2dec and cfg.nconds > 0:
        prop = linear(out[:, : cfg.nconds, :], sd, "prop_fc")
        mol = out[:, cfg.nconds:, :]
    elif cfg.nconds > 0 or cfg.is_vaetf:
        prop = torch.zeros(out.size(0), cfg.nconds, 1)
        mol = out
    else:
        prop, mol = None, out
    return prop, mol, mu, log_var, z


def forward_propagation(sd, cfg: ModelCfg, batch: Dict[str, Tensor], pad_id: int, eps=None):
    """Model/forward_propagation1.py:4-48 -- masks from the batch, trg = batch['trg'][:, :-1]."""
    trg_in = batch["trg"][:, :-1]
    has_c = cfg.model_type in ("pvaetf", "pscavaetf")
    sm = src_mask(batch["src"], pad_id, cfg.nconds if has_c else 0)
    tm = trg_mask(trg_in, pad_id, cfg.use_cond2dec, cfg.nconds if has_c else 0)
    return forward(sd, cfg, batch["src"], trg_in, sm, tm,
                   batch.get("econds") if has_c else None,
                   batch.get("dconds") if has_c else None, eps)


# ----------------------------------------------------------------------------------------------
# a15  loss                                                          Train/trainer1.py:14-30
# ----------------------------------------------------------------------------------------------
def kl_annealer(epoch: int, ini: float, inc: float, beg: int) -> float:
    return ini + inc * ((epoch + 1) - beg)


def loss_function(beta, preds_prop, preds_mol, ys_cond, ys_mol, mu, log_var, use_cond2dec, pad_id):
    """CE(sum, ignore pad) + beta*KL(sum over ALL positions) [+ MSE(sum) iff cond2dec]."""
    rce = F.cross_entropy(preds_mol.contiguous().view(-1, preds_mol.size(-1)), ys_mol,
                          ignore_index=pad_id, reduction="sum")
    kld = -0.5 * torch.sum(1 + log_var - mu.pow(2) - log_var.exp())
    if use_cond2dec:
        rprop = F.mse_loss(preds_prop, ys_cond, reduction="sum")
        loss = rce + rprop + beta * kld
    else:
        rprop = torch.zeros(1)
        loss = rce + beta * kld
    return loss, rce, rprop, kld


# ----------------------------------------------------------------------------------------------
# a16  optimiser step pieces                                         Train/trainer1.py:112-127
# ----------------------------------------------------------------------------------------------
def noam_lr(step: int, d_model: int, warmup: int) -> float:
    """d^-0.5 * min(step^-0.5, step*warm^-1.5) -- assigned AFTER optimizer.step()."""
    return float(d_model) ** -0.5 * min(float(step) ** -0.5, float(step) * float(warmup) ** -1.5)


def adam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float,
              b1: float = 0.9, b2: float = 0.98, eps: float = 1e-9) -> None:
    """torch.optim.Adam (no weight decay, no amsgrad) as configured at train1.py:116-119."""
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


# ----------------------------------------------------------------------------------------------
# a17  Sampling.decode (un-cached, exactly the reference's loop)     Inference/sampling_tool.py:140-184
# ----------------------------------------------------------------------------------------------
def sampling_decode(sd, cfg: ModelCfg, zs, ys, smask, dconds=None, *, max_strlen=80, pad_id=1,
                    eos_id=3, algo="greedy", uniforms: Optional[Tensor] = None) -> Tensor:
    """Re-runs the whole decoder on the prefix each step, keeps the last position's softmax,
    appends argmax (first max index) or a categorical draw; finished rows are NOT frozen; stops
    when every row has emitted <eos> at least once or after max_strlen-1 steps.

    `uniforms` (steps, n) replaces torch.multinomial with inverse-CDF sampling on supplied
    U(0,1) numbers so that a device implementation can be compared draw-for-draw."""
    done = torch.zeros(ys.size(0), dtype=torch.bool)
    nc = cfg.nconds if cfg.model_type in ("pvaetf", "pscavaetf") else 0
    for i in range(max_strlen - 1):
        tm = trg_mask(ys, pad_id, cfg.use_cond2dec, nc)
        logits = decode_logits(sd, cfg, ys, zs, smask, tm, dconds)
        if cfg.use_cond2dec:
            logits = logits[:, nc:, :]
        prob = torch.softmax(logits, dim=-1)[:, -1, :]
        if algo == "greedy":
            nxt = prob.max(dim=1)[1]
        elif uniforms is not None:
            nxt = inverse_cdf_draw(prob, uniforms[i])
        else:
            nxt = torch.multinomial(prob, 1).squeeze(-1)
        ys = torch.cat([ys, nxt.unsqueeze(-1)], dim=1)
        done |= (nxt.cpu() == eos_id)
        if bool(done.all()):
            break
    return ys


def inverse_cdf_draw(prob: Tensor, u: Tensor) -> Tensor:
    """Smallest index k with cumsum(prob)[k] > u*sum(prob) (sequential fp32 cumsum)."""
    c = torch.cumsum(prob.float(), dim=1)
    thr = (u.to(prob.device).float() * c[:, -1]).unsqueeze(1)
    idx = (c > thr).float().argmax(dim=1)
    none = ~(c > thr).any(dim=1)
    idx[none] = prob.size(1) - 1
    return idx


def id_to_smi(ids, itos, sos_id=2, eos_id=3) -> str:
    """Inference/sampling_tool.py:54-61."""
    out = ""
    for i in ids:
        if i == eos_id:
            break
        if i != sos_id:
            out += itos[int(i)]
    return out


# ----------------------------------------------------------------------------------------------
# a18  token-length sampler                                          Inference/toklen_sampling.py:4-36
# ----------------------------------------------------------------------------------------------
def toklen_from_distribution(data: np.ndarray, size: int, n_bins: int) -> np.ndarray:
    """Histogram-CDF sampler with a half-bin-width Gaussian jitter; consumes the global NumPy
    RNG in the order (uniform, normal) per draw, like the reference."""
    count, bins = np.histogram(data, bins=n_bins)
    pdf = count / np.sum(count)
    dx = np.diff(bins)[0]
    xc = bins[0:-1] + 0.5 * dx
    cdf = np.zeros_like(bins)
    cdf[1:] = np.cumsum(pdf)
    out = []
    for _ in range(size):
        a = np.random.uniform(0, 1)
        idx = np.argmax(cdf >= a) - 1
        out.append(xc[idx] + dx * np.random.normal() / 2)
    return np.array(out).reshape(size, 1)


def sample_toklen(toklen_data: np.ndarray, n: int, cond_dim: int) -> np.ndarray:
    """Inference/sampling_tool.py:75-81."""
    n_bin = int(toklen_data.max() - toklen_data.min())
    t = toklen_from_distribution(toklen_data, n, n_bin).reshape((-1,)) + cond_dim
    return np.rint(t).astype(int)


# ----------------------------------------------------------------------------------------------
# helpers for tests / bench (not part of the reference)
# ----------------------------------------------------------------------------------------------
def flops_forward(cfg: ModelCfg, B: int, S: int, T: int) -> float:
    """Algorithmic forward FLOPs, formula of SURVEY.md section 8(d)."""
    d, dff, lat, N = cfg.d_model, cfg.dff, cfg.latent_dim, cfg.N
    nc = cfg.nconds
    Se = nc + S
    Sm = Se + (nc if cfg.use_cond2lat else 0)
    Ne, Nd, Nm = B * Se, B * T, B * Sm
    enc = N * (2 * Ne * (4 * d * d + 2 * d * dff) + 4 * B * Se * Se * d)
    dec = N * (2 * Nd * 4 * d * d + 2 * Nd * 2 * d * d + 2 * Nm * 2 * d * d + 2 * Nd * 2 * d * dff
               + 4 * B * T * T * d + 4 * B * T * Sm * d)
    heads = 2 * Ne * d * 2 * lat + 2 * Ne * lat * d + 2 * Nd * d * cfg.trg_vocab
    return float(enc + dec + heads)
