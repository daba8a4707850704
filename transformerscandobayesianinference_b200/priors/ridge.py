"""`priors.ridge` (reference priors/ridge.py:12-18): noisy linear functions with N(0, .1^2) weights."""
import torch

from .utils import get_batch_to_dataloader


def get_batch(batch_size, seq_len, num_features, noisy_std=.1, device='cpu'):
    m = torch.normal(0., .1, size=(batch_size, num_features), device=device)
    x = torch.rand(seq_len, batch_size, num_features, device=device)
    y_non_noisy = torch.einsum('bf,tbf->tb', m, x)
    y = y_non_noisy + torch.normal(torch.zeros_like(y_non_noisy), noisy_std)
    return x, y, y_non_noisy


DataLoader = get_batch_to_dataloader(get_batch)
DataLoader.num_outputs = 1
