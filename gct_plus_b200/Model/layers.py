"""Containers for Model/layers.py of the reference (parameters only; see modules.py)."""
import torch.nn as nn

from .modules import Norm
from .sublayers import FeedForward, MultiHeadAttention


class EncoderLayer(nn.Module):
    def __init__(self, heads, d_model, dff, dropout, get_attn=False):
        super().__init__()
        self.get_attn = get_attn
        self.norm_1 = Norm(d_model)
        self.attn = MultiHeadAttention(heads, d_model, dropout, get_attn)
        self.dropout_1 = nn.Dropout(dropout)
        self.norm_2 = Norm(d_model)
        self.ff = FeedForward(d_model, dff, dropout)
        self.dropout_2 = nn.Dropout(dropout)


class DecoderLayer(nn.Module):
    def __init__(self, heads, d_model, dff, dropout, get_attn=False):
        super().__init__()
        self.get_attn = get_attn
        self.norm_1 = Norm(d_model)
        self.attn_1 = MultiHeadAttention(heads, d_model, dropout, get_attn)
        self.dropout_1 = nn.Dropout(dropout)
        self.norm_2 = Norm(d_model)
        self.attn_2 = MultiHeadAttention(heads, d_model, dropout, get_attn)
        self.dropout_2 = nn.Dropout(dropout)
        self.norm_3 = Norm(d_model)
        self.ff = FeedForward(d_model, dff, dropout)
        self.dropout_3 = nn.Dropout(dropout)
