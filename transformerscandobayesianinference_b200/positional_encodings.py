"""Positional encodings with the reference's protocol (reference positional_encodings.py):
`__init__(d_model, max_len)`, `forward(x [T, B, d_model]) -> same shape`.  PFN training on permutation-invariant
priors uses `NoPositionalEncoding` (reference train.py:42), which the fused embedding kernel folds away."""
import math

import torch
from torch import nn


class NoPositionalEncoding(nn.Module):
    def __init__(self, d_model, max_len=None):
        super().__init__()

    def forward(self, x):
        return x


class PositionalEncoding(nn.Module):
    """Fixed sinusoidal table (reference positional_encodings.py:21-34)."""

    def __init__(self, d_model, max_len=5000):
        super().__init__()
        position = torch.arange(max_len, dtype=torch.float).unsqueeze(1)
        freq = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
        table = torch.zeros(max_len, 1, d_model)
        table[:, 0, 0::2] = torch.sin(position * freq)
        table[:, 0, 1::2] = torch.cos(position * freq)
        self.register_buffer('pe', table)

    def forward(self, x):
        return x + self.pe[:x.size(0)]


class LearnedPositionalEncoding(nn.Module):
    """Trainable table, N(0, 1/d_model) init (reference positional_encodings.py:37-49)."""

    def __init__(self, d_model, max_len=5000):
        super().__init__()
        self.max_seq_len = max_len
        self.positional_embeddings = nn.Parameter(torch.empty(max_len, d_model))
        nn.init.normal_(self.positional_embeddings, mean=0, std=d_model ** -0.5)

    def forward(self, x):
        seq_len = x.shape[0]
        assert seq_len <= len(self.positional_embeddings), 'seq_len can be at most max_len.'
        return x + self.positional_embeddings[:seq_len].unsqueeze(1)


class PairedScrambledPositionalEncodings(LearnedPositionalEncoding):
    """Learned table whose (pair-grouped) rows are randomly permuted on every call (reference :52-62)."""

    def forward(self, x):
        seq_len = x.shape[0]
        table = self.positional_embeddings
        assert seq_len <= len(table), 'seq_len can be at most max_len.'
        assert len(table) % 2 == 0, 'Please specify an even max_len.'
        pairs = table.view(len(table), -1, 2)
        scrambled = pairs[torch.randperm(len(pairs))].view(*table.shape)[:seq_len]
        return x + scrambled.unsqueeze(1)
