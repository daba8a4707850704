"""Fit of the logistic-form GELU used by the tensor-core GEMM epilogues (csrc/common.cuh: gelu_fast / gelu_grad_fast).

Phi(u) ~= 1 / (1 + exp(-2 u (c0 + c1 u^2 + c2 u^4))); minimises max(|dGELU|, |dGELU'| / 2) against the exact erf form."""
import numpy as np
from scipy.special import erf
from scipy.optimize import minimize

u = np.linspace(-9, 9, 180001)
Phi = 0.5 * (1 + erf(u / np.sqrt(2)))
phi = np.exp(-0.5 * u * u) / np.sqrt(2 * np.pi)
G, dG = u * Phi, Phi + u * phi


def errs(c):
    u2 = u * u
    q = c[0] + u2 * (c[1] + u2 * c[2])
    dp = c[0] + u2 * (3 * c[1] + u2 * 5 * c[2])
    s = 1 / (1 + np.exp(-2 * u * q))
    d = s + u * s * (1 - s) * 2 * dp
    return np.max(np.abs(s - Phi)), np.max(np.abs(u * s - G)), np.max(np.abs(d - dG))


if __name__ == "__main__":
    c0 = [0.7978845608, 0.0356774, 0.0]
    r = minimize(lambda c: max(errs(c)[1], 0.5 * errs(c)[2]), c0, method="Nelder-Mead",
                 options=dict(xatol=1e-13, fatol=1e-15, maxiter=40000))
    print("coefficients", list(r.x))
    print("max |dPhi|, |dGELU|, |dGELU'|:", errs(r.x))
