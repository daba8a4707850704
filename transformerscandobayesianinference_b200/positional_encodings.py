"""Positional encodings behind the reference's protocol (reference positional_encodings.py:8-10): a class is built as
`cls(d_model, max_len)` and called on `x [T, B, d_model]`, returning a tensor of the same shape.

All four variants are "add the first T rows of a [max_len, d_model] table, broadcast over the batch"; they differ only in
where the table comes from, so they share one base class here.  PFN training on permutation-invariant priors uses
`NoPositionalEncoding` (reference train.py:42); `TransformerModel` recognises that class and folds it away in front of the
fused embedding kernel, every other variant runs as a regular PyTorch module in front of the CUDA encoder stack.
State-dict keys (`pe`, `positional_embeddings`) and the parameter initialisation order are the reference's, so its
checkpoints load strictly and a seeded construction reproduces its initial weights."""
import math

import torch
from torch import nn


class _AddTableRows(nn.Module):
    """x + table[:T] broadcast over the batch dimension; subclasses provide `_rows(T)` -> [T, d_model]."""

    def _rows(self, n):
        raise NotImplementedError

    def forward(self, x):
        return x + self._rows(x.shape[0]).unsqueeze(1)


class NoPositionalEncoding(nn.Module):
    """Identity (the constructor takes and ignores the protocol's arguments)."""

    def __init__(self, d_model, max_len=None):
        super().__init__()

    def forward(self, x):
        return x


class PositionalEncoding(_AddTableRows):
    """Fixed sinusoids: column 2i = sin(t w_i), column 2i+1 = cos(t w_i), w_i = 10000^(-2i/d)  (reference :21-34).
    The buffer keeps the reference's name and [max_len, 1, d_model] shape."""

    def __init__(self, d_model, max_len=5000):
        super().__init__()
        rates = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
        angles = torch.arange(max_len, dtype=torch.float)[:, None] * rates[None, :]
        table = torch.zeros(max_len, 1, d_model)
        table[:, 0, 0::2], table[:, 0, 1::2] = angles.sin(), angles.cos()
        self.register_buffer('pe', table)

    def _rows(self, n):
        return self.pe[:n, 0]


class LearnedPositionalEncoding(_AddTableRows):
    """A trainable table drawn from N(0, 1/d_model)  (reference :37-49)."""

    def __init__(self, d_model, max_len=5000):
        super().__init__()
        self.max_seq_len = max_len
        self.positional_embeddings = nn.Parameter(torch.empty(max_len, d_model))
        nn.init.normal_(self.positional_embeddings, mean=0, std=d_model ** -0.5)

    def _check(self, n):
        assert n <= len(self.positional_embeddings), 'seq_len can be at most max_len.'

    def _rows(self, n):
        self._check(n)
        return self.positional_embeddings[:n]


class PairedScrambledPositionalEncodings(LearnedPositionalEncoding):
    """The learned table with its rows re-ordered on every call: the flat table is regrouped into d_model/2 blocks of
    [max_len, 2], the blocks' leading index is permuted and the result cut to T rows (reference :52-62; one permutation per
    call, shared by the whole batch; one `torch.randperm` draw per call as in the reference)."""

    def _rows(self, n):
        self._check(n)
        table = self.positional_embeddings
        assert len(table) % 2 == 0, 'Please specify an even max_len.'
        blocks = table.view(len(table), -1, 2)
        return blocks[torch.randperm(len(blocks))].view(*table.shape)[:n]
