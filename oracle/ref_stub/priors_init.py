"""Replacement `priors/__init__.py` for the vendored reference tree (oracle/_ref): the reference's own init imports
gp / pyro / stroke / omniglot priors that need gpytorch, botorch, pyro and datasets (priors/__init__.py:1).
Only the priors of the hot path are exposed.  TEST / BASELINE INFRASTRUCTURE ONLY."""
from . import fast_gp, mlp, utils, prior  # noqa: F401
