"""Drop-in `bar_distribution` module (reference bar_distribution.py): the bucketised "Riemann" density head.

`forward(logits, y)` — the training criterion — runs as one fused CUDA kernel forward (bucket lookup + online
log-sum-exp + gather + width scaling + half-normal tails) and one backward (softmax - onehot), see
csrc/bar_nll.cu.  The inference helpers (`mean`, `quantile`, `mode`, `ei`) are small dense tensor expressions
and stay in PyTorch (they are not on the training hot path).
"""
import torch
from torch import nn

from . import _lib as L

_ICDF_HALFNORMAL_HALF = 0.6744897501960817  # HalfNormal(1).icdf(0.5)


class _BarNLLFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, y, borders, n_bars, full_support, oob_count):
        L.require_cuda(logits, y, borders)
        lead_shape = logits.shape[:-1]
        lg = logits.detach().reshape(-1, n_bars)
        if lg.dtype not in (torch.float32, torch.bfloat16):
            lg = lg.float()
        if lg.stride(-1) != 1:
            lg = lg.contiguous()
        yv = y.detach().reshape(-1).float().contiguous()
        assert yv.numel() == lg.shape[0], f'{yv.numel()} targets for {lg.shape[0]} logit rows'
        rows = lg.shape[0]
        dev = lg.device
        nll = torch.empty(rows, device=dev, dtype=torch.float32)
        idx = torch.empty(rows, device=dev, dtype=torch.int64)
        lse = torch.empty(rows, device=dev, dtype=torch.float32)
        if rows:
            L.bar_nll_fwd(lg, yv, borders, n_bars, full_support, nll, idx, lse, oob_count)
        ctx.save_for_backward(lg, idx, lse)
        ctx.meta = (n_bars, logits.shape, logits.dtype)
        return nll.reshape(lead_shape)

    @staticmethod
    def backward(ctx, g):
        lg, idx, lse = ctx.saved_tensors
        n_bars, shape, dtype = ctx.meta
        rows = lg.shape[0]
        dl = torch.empty(rows, n_bars, device=lg.device, dtype=torch.float32)
        if rows:
            L.bar_nll_bwd(lg, idx, lse, g.reshape(-1).float().contiguous(), dl, n_bars)
        return dl.reshape(shape).to(dtype), None, None, None, None, None


class BarDistribution(nn.Module):
    """borders: sorted 1-D tensor, first = min, last = max of the support (reference bar_distribution.py:5-17)."""

    # When True the in-support assertion of `forward` (reference :27, a device->host sync per call) is deferred:
    # violations are only counted on the device and `check_support()` raises later.  `train()` turns this on and
    # checks once per epoch.
    defer_support_check = False

    def __init__(self, borders: torch.Tensor):
        super().__init__()
        assert len(borders.shape) == 1
        self.register_buffer('borders', borders)
        self.register_buffer('bucket_widths', self.borders[1:] - self.borders[:-1])
        full_width = self.bucket_widths.sum()
        assert (full_width - (self.borders[-1] - self.borders[0])).abs() < 1e-4, \
            f'diff: {full_width - (self.borders[-1] - self.borders[0])}'
        assert (torch.argsort(borders) == torch.arange(len(borders))).all(), "Please provide sorted borders!"
        self.num_bars = len(borders) - 1
        self._oob = None

    _full_support = False

    def _oob_counter(self, device):
        if self._oob is None or self._oob.device != device:
            self._oob = torch.zeros(1, device=device, dtype=torch.int32)
        return self._oob

    def check_support(self):
        """Raise if any target seen since the last check fell outside (min, max) — the reference's assert (:27)."""
        if self._oob is not None:
            bad = int(self._oob.item())
            self._oob.zero_()
            assert bad == 0, f'{bad} targets y not in support set for borders (min_y, max_y) {self.borders}'

    def map_to_bucket_idx(self, y):
        """Bucket k = (b_k, b_{k+1}], left edge of bucket 0 inclusive; out-of-range -> -1 / num_bars (:19-23)."""
        L.require_cuda(y, self.borders)
        yv = y.detach().reshape(-1).float().contiguous()
        idx = torch.empty(yv.numel(), device=y.device, dtype=torch.int64)
        if yv.numel():
            L.bar_bucket_idx(yv, self.borders, self.num_bars, idx)
        return idx.reshape(y.shape)

    def forward(self, logits, y):
        """Negative log density of y under the bar distribution given by `logits` (… x num_bars), y: (…)."""
        assert logits.shape[-1] == self.num_bars, f'{logits.shape[-1]} vs {self.num_bars}'
        if self._full_support:
            assert self.num_bars > 1
        borders = self.borders if self.borders.dtype == torch.float32 else self.borders.float()
        nll = _BarNLLFn.apply(logits, y, borders, self.num_bars, self._full_support, self._oob_counter(logits.device))
        if not self._full_support and not BarDistribution.defer_support_check:
            self.check_support()
        return nll

    # ---- inference helpers (reference :35-80), plain tensor expressions ---------------------------------
    def _bucket_means(self):
        return self.borders[:-1] + self.bucket_widths / 2

    def mean(self, logits):
        return torch.softmax(logits.float(), -1) @ self._bucket_means()

    def quantile(self, logits, center_prob=.682):
        """Lower / upper quantile of the central `center_prob` mass by linear interpolation inside the bucket that
        crosses it (reference :40-63, vectorised instead of a per-row Python loop)."""
        shape = logits.shape
        probs = logits.float().reshape(-1, shape[-1]).softmax(-1)
        side = (1 - center_prob) / 2
        borders = self.borders.float()

        def lower(pr, bd):
            cum = torch.cumsum(pr, -1)
            tgt = torch.full((pr.shape[0], 1), side, device=pr.device, dtype=pr.dtype)
            idx = torch.searchsorted(cum, tgt).clamp(0, cum.shape[-1] - 1).squeeze(-1)
            rows = torch.arange(pr.shape[0], device=pr.device)
            left_prob = cum[rows, idx - 1]          # idx == 0 wraps to the last entry exactly like the reference
            rest = side - left_prob
            left, right = bd[idx], bd[idx + 1]
            return left + (right - left) * rest / pr[rows, idx]

        lo = lower(probs, borders)
        hi = lower(probs.flip(-1), borders.flip(0))
        return torch.stack([lo, hi], -1).reshape(*shape[:-1], 2)

    def mode(self, logits):
        return self._bucket_means()[logits.argmax(-1)]

    def ei(self, logits, best_f, maximize=True):
        """Expected improvement over `best_f` with each bucket treated as uniform (reference :70-80)."""
        lo, hi = self.borders[:-1].float(), self.borders[1:].float()
        best = torch.as_tensor(best_f, dtype=lo.dtype, device=lo.device)
        if maximize:
            contrib = ((hi + torch.maximum(lo, best)) / 2 - best).clamp(min=0)
        else:
            contrib = -((torch.minimum(hi, best) + lo) / 2 - best).clamp(max=0)
        return torch.softmax(logits.float(), -1) @ contrib.to(logits.device)


class FullSupportBarDistribution(BarDistribution):
    """Bar distribution whose first / last bucket are replaced by half-normal tails (reference :83-117)."""
    _full_support = True

    @staticmethod
    def halfnormal_with_p_weight_before(range_max, p=.5):
        s = range_max / torch.distributions.HalfNormal(torch.tensor(1.)).icdf(torch.tensor(p))
        return torch.distributions.HalfNormal(s)

    def mean(self, logits):
        means = self._bucket_means().clone()
        s0 = self.bucket_widths[0] / _ICDF_HALFNORMAL_HALF
        s1 = self.bucket_widths[-1] / _ICDF_HALFNORMAL_HALF
        hn_mean = (2.0 / torch.pi) ** 0.5
        means[0] = self.borders[1] - s0 * hn_mean
        means[-1] = self.borders[-2] + s1 * hn_mean
        return torch.softmax(logits.float(), -1) @ means


def get_bucket_limits(num_outputs: int, full_range: tuple = None, ys: torch.Tensor = None):
    """Borders of `num_outputs` buckets: uniform over `full_range`, or equal-mass from a sample `ys` with the
    limits placed half-way between neighbouring chunk ends (reference bar_distribution.py:121-143)."""
    assert (ys is not None) or (full_range is not None)
    if ys is not None:
        ys = ys.flatten()
        cut = len(ys) % num_outputs
        if cut:
            ys = ys[:-cut]
        print(f'Using {len(ys)} y evals to estimate {num_outputs} buckets. Cut off the last {cut} ys.')
        per_bucket = len(ys) // num_outputs
        if full_range is None:
            full_range = (ys.min(), ys.max())
        else:
            assert full_range[0] <= ys.min() and full_range[1] >= ys.max()
            full_range = torch.tensor(full_range)
        ys_sorted = ys.sort(0).values
        inner = (ys_sorted[per_bucket - 1::per_bucket][:-1] + ys_sorted[per_bucket::per_bucket]) / 2
        print(full_range)
        limits = torch.cat([full_range[0].unsqueeze(0), inner, full_range[1].unsqueeze(0)], 0)
    else:
        width = (full_range[1] - full_range[0]) / num_outputs
        limits = torch.cat([full_range[0] + torch.arange(num_outputs).float() * width,
                            torch.tensor(full_range[1]).unsqueeze(0)], 0)
    assert len(limits) - 1 == num_outputs and full_range[0] == limits[0] and full_range[-1] == limits[-1]
    return limits
