"""Prior -> DataLoader adapter and small sampling helpers (reference priors/utils.py).  No plotting imports: the
reference pulls in matplotlib at module import (priors/utils.py:10-11), which is not part of the hot path."""
import random

import numpy as np
import scipy.stats as stats
import torch
from torch import nn

from ..utils import set_locals_in_self
from .prior import PriorDataLoader


def get_batch_to_dataloader(get_batch_method_):
    """Wrap a `get_batch(batch_size, seq_len, num_features, ...) -> (x, y, target_y)` function into a loader class
    that yields `num_steps` freshly sampled batches per epoch as `((x, y), target_y)` (reference :14-42)."""

    class DL(PriorDataLoader):
        get_batch_method = get_batch_method_

        # `num_features` may be a class attribute set before instantiation when it is not part of the kwargs.
        def __init__(self, num_steps, fuse_x_y=False, **get_batch_kwargs):
            set_locals_in_self(locals())
            self.num_features = get_batch_kwargs.get('num_features') or self.num_features
            self.num_outputs = get_batch_kwargs.get('num_outputs') or self.num_outputs
            print('DataLoader.__dict__', self.__dict__)

        @staticmethod
        def gbm(*args, fuse_x_y=True, **kwargs):
            x, y, target_y = get_batch_method_(*args, **kwargs)
            if fuse_x_y:
                shifted = torch.cat([torch.zeros_like(y[:1]), y[:-1]], 0).unsqueeze(-1).float()
                return torch.cat([x, shifted], -1), target_y
            return (x, y), target_y

        def __len__(self):
            return self.num_steps

        def __iter__(self):
            return iter(self.gbm(**self.get_batch_kwargs, fuse_x_y=self.fuse_x_y) for _ in range(self.num_steps))

    return DL


trunc_norm_sampler_f = lambda mu, sigma: lambda: stats.truncnorm((0 - mu) / sigma, (1 - mu) / sigma, loc=mu, scale=sigma).rvs(1)[0]
beta_sampler_f = lambda a, b: lambda: np.random.beta(a, b)
gamma_sampler_f = lambda a, b: lambda: np.random.gamma(a, b)
uniform_sampler_f = lambda a, b: lambda: np.random.uniform(a, b)
uniform_int_sampler_f = lambda a, b: lambda: np.random.randint(a, b)
zipf_sampler_f = lambda a, b, c: lambda: min(b + np.random.zipf(a), c)
scaled_beta_sampler_f = lambda a, b, scale, minimum: lambda: minimum + round(beta_sampler_f(a, b)() * (scale - minimum + 1) - 0.5)


def normalize_data(data):
    """Zero mean / unit (unbiased) std over dim 0, eps 1e-6 (reference :73-78)."""
    return (data - data.mean(0)) / (data.std(0) + .000001)


def normalize_by_used_features_f(x, num_features_used, num_features):
    return x / (num_features_used / num_features)


class Binarize(nn.Module):
    """1 where x exceeds the (lower) median of the WHOLE tensor (reference :85-91)."""

    def __init__(self, p=0.5):
        super().__init__()
        self.p = p

    def forward(self, x):
        return (x > torch.median(x)).float()


def order_by_y(x, y):
    """Sort the sequence by (+/-) y of the first dataset, then interleave the two halves (reference :94-100)."""
    order = torch.argsort(y if random.randint(0, 1) else -y, dim=0)[:, 0, 0]
    order = order.reshape(2, -1).transpose(0, 1).reshape(-1)
    return x[order], y[order]
