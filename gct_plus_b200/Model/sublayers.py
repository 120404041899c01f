"""Containers for Model/sublayers.py of the reference (parameters only; see modules.py)."""
import torch.nn as nn


class Sampler(nn.Module):
    """mu / log_var heads of Vaetf (reference Model/sublayers.py:7-26)."""

    def __init__(self, d_model, latent_dim, variational):
        super().__init__()
        self.variational = variational
        self.fc_mu = nn.Linear(d_model, latent_dim)
        self.fc_log_var = nn.Linear(d_model, latent_dim)


class MultiHeadAttention(nn.Module):
    """Registration order q, v, k, out fixes the init RNG order (reference sublayers.py:54-59)."""

    def __init__(self, heads, d_model, dropout=0.1, get_attn=False):
        super().__init__()
        self.d_model = d_model
        self.d_k = d_model // heads
        self.h = heads
        self.get_attn = get_attn
        self.q_linear = nn.Linear(d_model, d_model)
        self.v_linear = nn.Linear(d_model, d_model)
        self.k_linear = nn.Linear(d_model, d_model)
        self.dropout = nn.Dropout(dropout)
        self.out = nn.Linear(d_model, d_model)


class FeedForward(nn.Module):
    def __init__(self, d_model, d_ff=2048, dropout=0.1):
        super().__init__()
        self.linear_1 = nn.Linear(d_model, d_ff)
        self.dropout = nn.Dropout(dropout)
        self.linear_2 = nn.Linear(d_ff, d_model)
