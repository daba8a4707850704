"""Robustness of the GP sampler's Cholesky at the cfg-2 shape: how many of the datasets report a failing pivot at each jitter
level (the loader retries those datasets with the next jitter; a dataset failing at 1e-4 raises NotPSDError)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transformerscandobayesianinference_b200 import _lib as L
dev = torch.device("cuda:0")
Bn, T = 512, 1000
ls = torch.full((Bn, 1), .6, device=dev); os_ = torch.ones(Bn, device=dev); nz = torch.full((Bn,), 1e-4, device=dev)
y = torch.empty(Bn, T, device=dev); work = torch.empty(Bn, T, T, device=dev)
for seed in range(8):
    torch.manual_seed(seed)
    x = torch.rand(Bn, T, 1, device=dev); z = torch.randn(Bn, T, device=dev)
    row = []
    for jit in (0.0, 1e-6, 1e-5, 1e-4):
        info = torch.zeros(Bn, device=dev, dtype=torch.int32)
        L.gp_sample(x, z, ls, os_, nz, jit, 0, y, work, info)
        torch.cuda.synchronize()
        bad = info != 0
        row.append(f"jit {jit:g}: {int(bad.sum())} bad" + (f" (first pivot {int(info[bad].min())})" if bad.any() else ""))
    print(f"seed {seed}: " + "; ".join(row), "| y finite:", bool(torch.isfinite(y).all()))
