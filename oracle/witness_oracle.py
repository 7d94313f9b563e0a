"""TEST INFRASTRUCTURE ONLY -- differentiable CPU oracle of the witness block ``kernel(X, Y, K_XY_only=True)``.

torch (CPU, fp64 by default) restatement of the K_XY branch of the reference kernels, gan/core/mmd.py:
``_distance_kernel`` :18-37 (mysqrt :12), ``_dot_kernel`` :44-52, ``_rbf_kernel`` :55-82, ``_mix_rbf_kernel``
:85-116, ``_mix_rq_kernel`` :143-188 and the tanh wrappers :40,139 -- written with torch ops so that autograd gives
first AND second derivatives (the gradient penalty, gan/core/model.py:327-350, differentiates the witness twice).
Pinned in tests/test_witness_cpu.py against the ``kxy|...`` fixtures of tests/golden/mmd_golden.npz, which were
produced by the reference itself (oracle/make_golden.py).  Only tests / smoke() may import this module.
"""
from __future__ import annotations

import torch

EPS = 1e-5   # mmd.py:6


def _sqdist(X, Y):
    """mmd.py:64-67 style: |x|^2 + |y|^2 - 2 x.y (Gram form, unclamped)."""
    return (X * X).sum(1)[:, None] + (Y * Y).sum(1)[None, :] - 2.0 * X @ Y.T


def kernel_xy(name, X, Y, sigma=1.0, wt=1.0, sigmas=(2.0, 5.0, 10.0, 20.0, 40.0, 80.0), alphas=(0.1, 1.0, 10.0),
              wts=None, add_dot=None):
    """K_XY [m, n] of kernel `name` (same names as the reference's ``_<name>_kernel``)."""
    if name.startswith("tanh_"):
        return kernel_xy(name[5:], torch.tanh(X), torch.tanh(Y), sigma, wt, sigmas, alphas, wts, add_dot)
    if name == "dot":
        return X @ Y.T
    if name == "distance":
        nx, ny = (X * X).sum(1), (Y * Y).sum(1)
        root = lambda v: torch.sqrt(torch.clamp(v + EPS, min=0.0))
        return root(nx)[:, None] + root(ny)[None, :] - root(_sqdist(X, Y))
    D = torch.clamp(_sqdist(X, Y), min=0.0)
    if name == "rbf":
        return wt * torch.exp(-D / (2.0 * sigma ** 2))
    if name == "mix_rbf":
        wts = [1.0] * len(sigmas) if wts is None else wts
        return sum(w * torch.exp(-D / (2.0 * s ** 2)) for s, w in zip(sigmas, wts))
    dots = {"mix_rq": 0.0, "mix_rq_dot": 0.1, "mix_rq_1dot": 1.0, "mix_rq_10dot": 10.0, "mix_rq_01dot": 0.1,
            "mix_rq_001dot": 0.01}
    if name in dots:
        c = dots[name] if add_dot is None else add_dot
        wts = [1.0] * len(alphas) if wts is None else wts
        K = sum(w * torch.exp(-a * torch.log(1.0 + D / (2.0 * a))) for a, w in zip(alphas, wts))
        return K + c * (X @ Y.T) if c > 0 else K
    raise ValueError("unknown kernel %r" % name)


def witness(name, x_hat, real, fake, **kw):
    """model.py:336-338: E_y k(x_hat, real) - E_y k(x_hat, fake), one value per row of x_hat."""
    return kernel_xy(name, x_hat, real, **kw).mean(1) - kernel_xy(name, x_hat, fake, **kw).mean(1)
