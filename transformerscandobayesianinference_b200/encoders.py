"""Input encoders with the reference's names (reference encoders.py).  `Linear` feeds the fused embedding kernel;
anything else runs as a regular PyTorch module in front of the CUDA encoder stack."""
import math

import torch
import torch.nn as nn

Linear = nn.Linear


class Normalize(nn.Module):
    """(x - mean) / std.  The reference refers to this class without defining it (encoders.py:18); provided so that
    `get_normalized_uniform_encoder` is usable."""

    def __init__(self, mean, std):
        super().__init__()
        self.mean, self.std = mean, std

    def forward(self, x):
        return (x - self.mean) / self.std


def get_normalized_uniform_encoder(encoder_creator):
    """Wrap an encoder fed with U[0,1] features so that it sees zero-mean / unit-std inputs (reference encoders.py:10-18)."""
    return lambda in_dim, out_dim: nn.Sequential(Normalize(.5, math.sqrt(1 / 12)), encoder_creator(in_dim, out_dim))


class CanEmb(nn.Embedding):
    """Per-feature class embedding, concatenated over features (reference encoders.py:22-33)."""

    def __init__(self, num_features, num_embeddings: int, embedding_dim: int, *args, **kwargs):
        assert embedding_dim % num_features == 0
        super().__init__(num_embeddings, embedding_dim // num_features, *args, **kwargs)

    def forward(self, x):
        emb = super().forward(x)
        return emb.view(*emb.shape[:-2], -1)


def get_Canonical(num_classes):
    return lambda num_features, emsize: CanEmb(num_features, num_classes, emsize)
