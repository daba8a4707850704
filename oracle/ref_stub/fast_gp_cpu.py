"""`priors.fast_gp` for the vendored reference tree (oracle/_ref): the reference's module builds a gpytorch ExactGP
(priors/fast_gp.py:13-32) and samples `model(x)` under `gpytorch.settings.prior_mode` (:48-56); gpytorch is not
installed, so the same draw is restated in plain torch (oracle/pfn_oracle.py: dense RBF kernel + noise on the diagonal,
`torch.linalg.cholesky`, root times randn) on whatever `device` is asked for — CPU for the cpu_baseline / reference arm,
CUDA (cuSOLVER) for the eager-GPU baseline.  The loader class comes from the UNMODIFIED priors/utils.py.
TEST / BASELINE INFRASTRUCTURE ONLY."""
import torch

from utils import default_device
from .utils import get_batch_to_dataloader


@torch.no_grad()
def get_batch(batch_size, seq_len, num_features, device=default_device, hyperparameters=None, equidistant_x=False):
    if isinstance(hyperparameters, (tuple, list)):
        hyperparameters = {"noise": hyperparameters[0], "outputscale": hyperparameters[1], "lengthscale": hyperparameters[2]}
    elif hyperparameters is None:
        hyperparameters = {"noise": .1, "outputscale": .1, "lengthscale": .1}
    if equidistant_x:
        assert num_features == 1
        x = torch.linspace(0, 1., seq_len).tile(batch_size, 1).unsqueeze(-1).to(device)
    else:
        x = torch.rand(batch_size, seq_len, num_features, device=device)
    xs = x / float(hyperparameters["lengthscale"])
    d2 = (xs.unsqueeze(2) - xs.unsqueeze(1)).pow(2).sum(-1)
    K = float(hyperparameters["outputscale"]) * torch.exp(-0.5 * d2)
    K = K + float(hyperparameters["noise"]) * torch.eye(seq_len, device=device)
    L = torch.linalg.cholesky(K)
    sample = (L @ torch.randn(batch_size, seq_len, 1, device=device)).squeeze(-1).transpose(0, 1)
    return x.transpose(0, 1), sample, sample


DataLoader = get_batch_to_dataloader(get_batch)
DataLoader.num_outputs = 1
