"""`priors.fast_gp_mix` (reference priors/fast_gp_mix.py): a mixture-of-GPs prior.  Every dataset draws its own
hyperparameters from Gamma hyperpriors (lengthscale per input dim ~ Gamma(3, 6), outputscale ~ Gamma(.5, .15),
noise ~ Gamma(1.1, .05), rate parameterisation; reference :24-52 via botorch SingleTaskGP.pyro_sample_from_prior)
and is then sampled from a Matern-nu ARD GP (nu default 2.5) by the same fused CUDA kernel as priors.fast_gp.

The botorch/pyro model objects the reference builds per group only serve to draw those hyperparameters; here they
are drawn directly with torch.distributions.Gamma on the device (botorch 0.6.0 / pyro 1.7.0 are not installed:
the hyper-prior restatement is validated distributionally — parity unpinned, see oracle/pfn_oracle.py).
"""
import random

import torch
from torch import nn

from .. import _lib as L
from ..bar_distribution import BarDistribution
from ..utils import default_device
from .fast_gp import _compute_device, sample_gp
from .utils import get_batch_to_dataloader, _Deferred

MIN_INFERRED_NOISE_LEVEL = 1e-4  # botorch.models.gp_regression.MIN_INFERRED_NOISE_LEVEL (noise constraint lower bound)

_NU_TO_KERNEL = {0.5: L.KERNEL_MATERN12, 1.5: L.KERNEL_MATERN32, 2.5: L.KERNEL_MATERN52}


def sample_hyperparameters(n, num_features, hyperparameters, device):
    """Per-dataset (lengthscale [n,F], outputscale [n], noise [n]) from the Gamma hyperpriors (reference :24-52)."""
    g = torch.distributions.Gamma
    hp = hyperparameters
    one = torch.ones((), device=device)
    ls = g(one * hp.get('lengthscale_concentration', 3.0), one * hp.get('lengthscale_rate', 6.0)).sample((n, num_features))
    os_ = g(one * hp.get('outputscale_concentration', .5), one * hp.get('outputscale_rate', 0.15)).sample((n,))
    noise = g(one * hp.get('noise_concentration', 1.1), one * hp.get('noise_rate', 0.05)).sample((n,))
    return ls.float().clamp_min(1e-6), os_.float().clamp_min(1e-10), noise.float().clamp_min(MIN_INFERRED_NOISE_LEVEL)


@torch.no_grad()
def get_batch(batch_size, seq_len, num_features, device=default_device, hyperparameters=None,
              batch_size_per_gp_sample=None, num_outputs=1, fix_to_range=None, equidistant_x=False, x=None, z=None):
    """-> x [T,B,F], y [T,B], target_y [T,B] (reference :58-134).  `x` / `z`: optional caller-supplied uniform inputs
    [B,T,F] and normal draws [B,T] (e.g. pinned host memory), see priors.fast_gp.get_batch; only without fix_to_range."""
    assert num_outputs == 1
    hps = hyperparameters or {}
    dev = _compute_device(device)
    batch_size_per_gp_sample = (batch_size_per_gp_sample or max(batch_size // 10, 1))
    assert batch_size % batch_size_per_gp_sample == 0
    kernel_type = _NU_TO_KERNEL[float(hps.get('nu', 2.5))]
    mult = 2 ** (fix_to_range is not None)
    total = batch_size * mult
    cand = batch_size_per_gp_sample * mult
    given_z = None
    if x is not None or z is not None:
        assert fix_to_range is None, "caller-supplied x / z cannot be combined with the rejection loop of fix_to_range"
    if z is not None:
        given_z = z.to(dev, torch.float32, non_blocking=True).contiguous()
    if x is not None:
        assert x.shape == (total, seq_len, num_features)
        x = x.to(dev, torch.float32, non_blocking=True).contiguous()
    elif equidistant_x:
        assert num_features == 1
        x = torch.linspace(0, 1., seq_len, device=dev).view(1, seq_len, 1).repeat(total, 1, 1)
    else:
        x = torch.rand(total, seq_len, num_features, device=dev)

    # post-processing reads y right away, so the deferred (sync-free) pivot check of sample_gp only applies to plain draws
    plain = fix_to_range is None and not hps.get('y_minmax_norm') and not hps.get('sigmoid')

    def draw(xs):
        if not plain and _Deferred.active:
            _Deferred.active = False
            try:
                return draw(xs)
            finally:
                _Deferred.active = True
        n = xs.shape[0]
        ls, os_, noise = sample_hyperparameters(n, num_features, hps, dev)
        z = given_z if given_z is not None else torch.randn(n, seq_len, device=dev)
        s = sample_gp(xs.contiguous(), z, ls, os_, noise, kernel_type)   # [n, T]
        if hps.get('y_minmax_norm'):
            lo, hi = s.min(1, keepdim=True)[0], s.max(1, keepdim=True)[0]
            s = (s - lo) / (hi - lo)
        if hps.get('sigmoid'):
            s = s.sigmoid()
        return s

    if fix_to_range is None:
        sample = draw(x)                      # all groups in one launch: the groups are independent draws
    else:
        pieces = []
        throwaway = 0.
        for i in range(0, total, cand):
            tries = 0
            while True:
                s = draw(x[i:i + cand])
                ok = ~((s < fix_to_range[0]) | (s >= fix_to_range[1])).any(1)
                throwaway += float((~ok[:batch_size_per_gp_sample]).sum()) / batch_size_per_gp_sample
                if int(ok.sum()) >= batch_size_per_gp_sample:
                    break
                tries += 1
                if tries < 100:
                    print("Please change hyper-parameters (e.g. decrease outputscale_mean) it"
                          "seems like the range is set to tight for your hyper-parameters.")
            x[i:i + batch_size_per_gp_sample] = x[i:i + cand][ok][:batch_size_per_gp_sample]
            pieces.append(s[ok][:batch_size_per_gp_sample])
        if random.random() < .01:
            print('throwaway share', throwaway / (batch_size // batch_size_per_gp_sample))
        sample = torch.cat(pieces, 0)
        x = x.view(-1, batch_size, seq_len, num_features)[0]
    x_t, y_t = x.transpose(0, 1), sample.transpose(0, 1)
    assert x_t.shape[:2] == y_t.shape[:2]
    out_dev = torch.device(device)
    if out_dev.type != 'cuda':
        x_t, y_t = x_t.to(out_dev), y_t.to(out_dev)
    return x_t, y_t, y_t


class DataLoader(get_batch_to_dataloader(get_batch)):
    num_outputs = 1

    @torch.no_grad()
    def validate(self, model, step_size=1, start_pos=0):
        """MSE of the bar-distribution mean at the first query row for every eval position (reference :140-153)."""
        if isinstance(model.criterion, BarDistribution):
            (x, y), target_y = self.gbm(**self.get_batch_kwargs, fuse_x_y=self.fuse_x_y)
            dev = next(model.parameters()).device
            x, y, target_y = x.to(dev), y.to(dev), target_y.to(dev)
            model.eval()
            losses = []
            mse = nn.MSELoss()
            for eval_pos in range(start_pos, len(x), step_size):
                logits = model((x, y), single_eval_pos=eval_pos)
                means = model.criterion.mean(logits)
                losses.append(mse(means[0], target_y[eval_pos]))
            model.train()
            return torch.stack(losses)
        return 123.
