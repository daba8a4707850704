"""Prior samplers of the PFN hot path (reference priors/): fast_gp, fast_gp_mix, mlp (+ ridge as a tiny test prior).
Unlike the reference's `priors/__init__.py:1`, importing this package does not import gpytorch / botorch / pyro."""
from . import fast_gp, fast_gp_mix, mlp, ridge, utils, prior  # noqa: F401
