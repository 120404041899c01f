"""Unconditioned Transformer-VAE, drop-in for the reference's Model/vaetf.py (class Vaetf).

Same constructor, attributes, state_dict keys and return tuples; the arithmetic runs in
libgct_b200.so (see engine.py).  Note the reference's vaetf.Encoder owns fc_mu / fc_log_var that
its forward never uses (vaetf.py:26-27) -- they are kept so checkpoints load, and never read.
"""
import torch
import torch.nn as nn

from .engine import TransformerVAE
from .layers import DecoderLayer, EncoderLayer
from .modules import Embeddings, Norm, PositionalEncoding, get_clones
from .sublayers import Sampler


class Encoder(nn.Module):
    def __init__(self, vocab_size, d_model, N, h, dff, latent_dim, nconds, dropout, variational=True, get_attn=False):
        super().__init__()
        self.N, self.nconds, self.variational, self.get_attn = N, nconds, variational, get_attn
        self.embed_sentence = Embeddings(d_model, vocab_size)
        self.norm = Norm(d_model)
        self.pe = PositionalEncoding(d_model, dropout=dropout)
        self.layers = get_clones(EncoderLayer(h, d_model, dff, dropout, get_attn), N)
        self.fc_mu = nn.Linear(d_model, latent_dim)
        self.fc_log_var = nn.Linear(d_model, latent_dim)
        if nconds > 0:
            self.embed_cond2enc = nn.Linear(nconds, d_model * nconds)


class Decoder(nn.Module):
    def __init__(self, vocab_size, d_model, N, h, dff, latent_dim, nconds, dropout, use_cond2dec, use_cond2lat,
                 get_attn=False):
        super().__init__()
        self.N, self.nconds, self.d_model, self.get_attn = N, nconds, d_model, get_attn
        self.use_cond2dec, self.use_cond2lat = use_cond2dec, use_cond2lat
        self.embed = Embeddings(d_model, vocab_size)
        self.pe = PositionalEncoding(d_model, dropout=dropout)
        self.fc_z = nn.Linear(latent_dim, d_model)
        self.layers = get_clones(DecoderLayer(h, d_model, dff, dropout, get_attn), N)
        self.norm = Norm(d_model)
        if use_cond2dec and nconds > 0:
            self.embed_cond2dec = nn.Linear(nconds, d_model * nconds)
        if use_cond2lat and nconds > 0:
            self.embed_cond2lat = nn.Linear(nconds, d_model * nconds)


class Vaetf(TransformerVAE):
    def __init__(self, src_vocab, trg_vocab, N=6, d_model=256, dff=2048, h=8, latent_dim=64, dropout=0.1, nconds=3,
                 use_cond2dec=False, use_cond2lat=False, variational=True, get_attn=False, compute_dtype=None):
        super().__init__()
        self.nconds, self.get_attn = nconds, get_attn
        self.use_cond2dec, self.use_cond2lat = use_cond2dec, use_cond2lat
        self.encoder = Encoder(src_vocab, d_model, N, h, dff, latent_dim, nconds, dropout, variational, get_attn)
        self.decoder = Decoder(trg_vocab, d_model, N, h, dff, latent_dim, nconds, dropout, use_cond2dec, use_cond2lat,
                               get_attn)
        self.sampler = Sampler(d_model, latent_dim, variational)
        self.out = nn.Linear(d_model, trg_vocab)
        if use_cond2dec and nconds > 0:
            self.prop_fc = nn.Linear(trg_vocab, 1)
        self.reset_parameters()
        self._finalize(compute_dtype)

    def reset_parameters(self):
        for _, p in self.named_parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    def encode(self, src, src_mask, econds=None):
        _, mu, log_var, z, _ = self._run(src, None, src_mask, None, econds, None, run_decoder=False)
        return z, mu, log_var

    def decode(self, trg, z, src_mask, trg_mask, dconds=None):
        logits, *_ = self._run(None, trg, src_mask, trg_mask, None, dconds, run_encoder=False, z_in=z)
        return logits

    def forward(self, src, trg, src_mask, trg_mask, econds=None, dconds=None):
        output, mu, log_var, z, attn = self._run(src, trg, src_mask, trg_mask, econds, dconds, want_attn=self.get_attn)
        if self.use_cond2dec and self.nconds > 0:
            output_prop = torch.nn.functional.linear(output[:, :self.nconds, :], self.prop_fc.weight, self.prop_fc.bias)
            output_mol = output[:, self.nconds:, :]
        else:
            output_prop = torch.zeros(output.size(0), self.nconds, 1)
            output_mol = output
        if self.get_attn:
            return (output_prop, output_mol, mu, log_var, z, list(attn[0]), list(attn[1]), list(attn[2]))
        return output_prop, output_mol, mu, log_var, z
