"""TEST/BENCH INFRASTRUCTURE ONLY -- multi-threaded CPU port of the reference hot path, used as the
`cpu_baseline` / `--impl reference` arm of bench.py (the reference itself is TF-1.x Python that cannot
travel to the GPU box).

What the reference does on CPU, restated with torch-CPU fp32 (all host threads, like TF's Eigen pool):
  loss:  three matmuls -> squared norms from the Gram diagonal -> clamp -> elementwise kernel transform
         -> block sums (gan/core/mmd.py:143-188, 194-220), gradients by reverse-mode autodiff through the
         materialised N x N temporaries (gan/core/model.py:446,452 `tf.gradients`)
  KID:   per subset: fancy-index gather, three (XY^T/d + 1)^3 Gram blocks, _mmd2_and_variance sums
         (gan/compute_scores.py:211-335) -- numpy/BLAS exactly as the reference (via oracle/kid_oracle.py)
Pinned to the reference-minted golden fixtures and to the live reference in tests/test_oracle_cpu.py
(test_cpu_port_matches_reference_golden / _live).  bench.py uses it only when the reference sources are not staged under
oracle/_ref (then `kind` is "port").  Never imported by the product package.
"""
from __future__ import annotations

import numpy as np
import torch


def _blocks(X, Y):
    XX, XY, YY = X @ X.T, X @ Y.T, Y @ Y.T
    return XX, XY, YY, torch.diagonal(XX), torch.diagonal(YY)


def mix_rq_mmd2(X, Y, alphas=(0.1, 1.0, 10.0), wts=None, add_dot=0.0, biased=False):
    """mmd2(_mix_rq_kernel(X, Y)) on torch CPU tensors (differentiable)."""
    wts = [1.0] * len(alphas) if wts is None else wts
    XX, XY, YY, nx, ny = _blocks(X, Y)

    def kern(G, nr, nc):
        D = torch.clamp(-2.0 * G + nr[:, None] + nc[None, :], min=0.0)
        K = 0.0
        for a, w in zip(alphas, wts):
            K = K + w * torch.exp(-a * torch.log(1.0 + D / (2.0 * a)))
        if add_dot > 0:
            K = K + add_dot * G
        return K

    K_XX, K_XY, K_YY = kern(XX, nx, nx), kern(XY, nx, ny), kern(YY, ny, ny)
    m, n = float(X.shape[0]), float(Y.shape[0])
    if biased:
        return K_XX.sum() / (m * m) + K_YY.sum() / (n * n) - 2 * K_XY.sum() / (m * n)
    cd = float(sum(wts))
    return ((K_XX.sum() - m * cd) / (m * (m - 1)) + (K_YY.sum() - n * cd) / (n * (n - 1))
            - 2 * K_XY.sum() / (m * n))


def mix_rq_fwd_bwd(Xnp, Ynp, **kw):
    """One loss evaluation + gradients w.r.t. both feature sets; returns (value, dX, dY) numpy."""
    X = torch.from_numpy(Xnp).requires_grad_(True)
    Y = torch.from_numpy(Ynp).requires_grad_(True)
    v = mix_rq_mmd2(X, Y, **kw)
    gX, gY = torch.autograd.grad(v, [X, Y])
    return float(v.detach()), gX.numpy(), gY.numpy()


def kid_subsets(codes_g, codes_r, idx_g, idx_r, ret_var=False):
    """The reference's subset loop (numpy + BLAS) on explicit indices."""
    from . import kid_oracle

    m = min(len(codes_g), len(codes_r))
    return [kid_oracle.polynomial_mmd(codes_g[ig], codes_r[ir], var_at_m=m, ret_var=ret_var)
            for ig, ir in zip(idx_g, idx_r)]
