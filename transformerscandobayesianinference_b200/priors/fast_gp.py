"""`priors.fast_gp` (reference priors/fast_gp.py): one GP-prior draw per dataset,
    x ~ U[0,1)^(B x T x F),  y = chol(outputscale * RBF(x, x; lengthscale) + noise * I) z,  z ~ N(0, I),
sampled by the fused CUDA kernel csrc/gp_sampler.cu (kernel build + blocked fp32 Cholesky + L z in one pass)
instead of gpytorch's lazy kernel -> cuSOLVER potrf -> bmm chain (reference :13-32,48-56).

gpytorch semantics kept: default hyperparameters (:40), tuple/list hyperparameters (:37-38), the RNG order
`rand` (x) then `randn` (z), and psd_safe_cholesky's jitter escalation (1e-6, 1e-5, 1e-4 added to the whole
batch's diagonal, then NotPSD error).  `fast_computations` is accepted and ignored: the kernel always computes
the exact Cholesky root (what the notebook's `(False, False, False)` setting selects).
"""
import math

import torch

from .. import _lib as L
from ..utils import default_device
from .utils import get_batch_to_dataloader, _Deferred

_JITTERS = (0.0, 1e-6, 1e-5, 1e-4)


class NotPSDError(RuntimeError):
    pass


def _compute_device(device):
    dev = torch.device(device)
    if dev.type == 'cuda':
        return dev
    if not torch.cuda.is_available():
        raise RuntimeError("priors.fast_gp samples with the sm_100a GP kernel; no CUDA device is available "
                           "(there is no CPU fallback)")
    return torch.device('cuda', torch.cuda.current_device())


def sample_gp(x, z, lengthscale, outputscale, noise, kernel_type=L.KERNEL_RBF, return_factor=False):
    """x [B,T,F], z [B,T] on a CUDA device; lengthscale [B,F], outputscale [B], noise [B] -> y [B,T]."""
    Bn, T, F = x.shape
    dev = x.device
    ldw = (T + 3) // 4 * 4
    y = torch.empty(Bn, T, device=dev, dtype=torch.float32)
    work = torch.empty(Bn, T, ldw, device=dev, dtype=torch.float32)
    info = torch.empty(Bn, device=dev, dtype=torch.int32)

    def attempt(jitters):
        for jitter in jitters:
            L.gp_sample(x, z, lengthscale, outputscale, noise, jitter, kernel_type, y, work, info)
            if not bool(info.any().item()):
                return
        raise NotPSDError(f"kernel matrix not positive definite even with jitter {_JITTERS[-1]:g} "
                          f"(first failing pivots: {info[info > 0][:8].tolist()})")

    if _Deferred.active and not return_factor:
        # No host sync here: the pivot flags travel to pinned host memory behind the kernel and are looked at when the
        # batch is handed to the consumer (a step later).  Only then -- and only if a pivot failed -- is the whole batch
        # re-factored with gpytorch's jitter escalation, overwriting y in place (every view handed out stays valid).
        L.gp_sample(x, z, lengthscale, outputscale, noise, _JITTERS[0], kernel_type, y, work, info)
        bad = info.max().reshape(1)
        bad_host = torch.empty(1, dtype=torch.int32, pin_memory=True)
        bad_host.copy_(bad, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        del work

        def resolve():
            ev.synchronize()
            if int(bad_host[0]) != 0:
                nonlocal_work = torch.empty(Bn, T, ldw, device=dev, dtype=torch.float32)
                _retry(x, z, lengthscale, outputscale, noise, kernel_type, y, nonlocal_work, info)
        _Deferred.pending.append(resolve)
        return y
    attempt(_JITTERS)
    return (y, torch.tril(work[:, :, :T].transpose(1, 2))) if return_factor else y


def _retry(x, z, lengthscale, outputscale, noise, kernel_type, y, work, info):
    with torch.cuda.device(x.device):
        for jitter in _JITTERS[1:]:
            L.gp_sample(x, z, lengthscale, outputscale, noise, jitter, kernel_type, y, work, info)
            if not bool(info.any().item()):
                return
    raise NotPSDError(f"kernel matrix not positive definite even with jitter {_JITTERS[-1]:g} "
                      f"(first failing pivots: {info[info > 0][:8].tolist()})")


def _hps_to_dict(hyperparameters):
    if isinstance(hyperparameters, (tuple, list)):
        return {"noise": hyperparameters[0], "outputscale": hyperparameters[1], "lengthscale": hyperparameters[2]}
    if hyperparameters is None:
        return {"noise": .1, "outputscale": .1, "lengthscale": .1}
    return hyperparameters


@torch.no_grad()
def get_batch(batch_size, seq_len, num_features, device=default_device, hyperparameters=None, equidistant_x=False,
              x=None, z=None):
    """-> x [T,B,F], y [T,B], target_y [T,B] (= y) on `device` (reference :36-58).

    Extension (defaults keep the reference signature): `x` [B,T,F] ~ U[0,1) and `z` [B,T] ~ N(0,1) may be supplied by
    the caller — e.g. drawn on the host and kept in pinned memory — instead of being drawn on the device; they are
    copied to the device asynchronously."""
    hps = _hps_to_dict(hyperparameters)
    dev = _compute_device(device)
    if x is not None:
        assert x.shape == (batch_size, seq_len, num_features)
        x = x.to(dev, torch.float32, non_blocking=True).contiguous()
    elif equidistant_x:
        assert num_features == 1
        x = torch.linspace(0, 1., seq_len, device=dev).view(1, seq_len, 1).repeat(batch_size, 1, 1)
    else:
        x = torch.rand(batch_size, seq_len, num_features, device=dev)
    if z is not None:
        assert z.shape == (batch_size, seq_len)
        z = z.to(dev, torch.float32, non_blocking=True).contiguous()
    else:
        z = torch.randn(batch_size, seq_len, device=dev)
    ls = torch.full((batch_size, num_features), float(hps["lengthscale"]), device=dev)
    os_ = torch.full((batch_size,), float(hps["outputscale"]), device=dev)
    noise = torch.full((batch_size,), float(hps["noise"]), device=dev)
    y = sample_gp(x, z, ls, os_, noise, L.KERNEL_RBF)
    x_t, y_t = x.transpose(0, 1), y.transpose(0, 1)
    out_dev = torch.device(device)
    if out_dev.type != 'cuda':
        x_t, y_t = x_t.to(out_dev), y_t.to(out_dev)
    return x_t, y_t, y_t


DataLoader = get_batch_to_dataloader(get_batch)
DataLoader.num_outputs = 1


@torch.no_grad()
def evaluate(x, y, y_non_noisy, use_mse=False, hyperparameters={}, get_model_on_device=None, device=default_device,
             step_size=1, start_pos=0):
    """Exact-GP posterior baseline (reference :88-120): for each t, condition on rows < t and score row t with the
    Gaussian predictive NLL (or MSE).  Restated with batched torch.linalg Cholesky solves on `device` — an
    evaluation utility, not part of the training hot path.  Returns (all_losses [n_t, B], mean losses, seconds)."""
    import time
    start = time.time()
    hps = _hps_to_dict(hyperparameters if hyperparameters else None)
    dev = torch.device(device)
    xb = x.to(dev, torch.float64).transpose(0, 1)            # [B,T,F]
    yb = y.to(dev, torch.float64).transpose(0, 1)            # [B,T]
    T = xb.shape[1]
    d2 = ((xb.unsqueeze(2) - xb.unsqueeze(1)) / float(hps["lengthscale"])).pow(2).sum(-1)
    K = float(hps["outputscale"]) * torch.exp(-0.5 * d2)
    noise = float(hps["noise"])
    means_list = [.0] if start_pos == 0 else []
    all_losses = []
    for t in range(max(start_pos, 1), T, step_size):
        Ktt = K[:, :t, :t] + noise * torch.eye(t, dtype=K.dtype, device=dev)
        kst = K[:, :t, t]                                    # [B,t]
        chol = torch.linalg.cholesky(Ktt)
        alpha = torch.cholesky_solve(yb[:, :t].unsqueeze(-1), chol).squeeze(-1)
        v = torch.cholesky_solve(kst.unsqueeze(-1), chol).squeeze(-1)
        mean = (kst * alpha).sum(-1)
        var = K[:, t, t] + noise - (kst * v).sum(-1)
        if use_mse:
            ls = (mean - yb[:, t]) ** 2
        else:
            ls = 0.5 * (math.log(2 * math.pi) + torch.log(var) + (yb[:, t] - mean) ** 2 / var)
        means_list.append(ls.mean().item())
        all_losses.append(ls.float().flatten())
    return torch.stack(all_losses).to('cpu'), torch.tensor(means_list).to('cpu'), time.time() - start
