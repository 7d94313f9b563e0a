"""TEST INFRASTRUCTURE ONLY -- torch restatement (CPU, fp64 = truth) of the reference's SMMD scaling:

  * squared_norm_jacobian ........ gan/core/ops.py:228-233  (one gradient per output feature, as the reference does)
  * add_scaling .................. gan/core/model.py:366-403 (x_hat = d_images, x_hat_data = images; variants 'grad' /
                                   'value_and_grad'; scale = 1 / (sc * ... + 1))
  * SMMD.set_loss / apply_scaling  gan/core/smmd.py:10-23    (g_loss = mmd2(kernel(G, images)) * scale, d_loss = -g_loss)
  * the dense rbf / mix_rq kernels and mmd2 they sit on: gan/core/mmd.py:55-82, 143-188, 194-220 (same formulas as
    oracle/mmd_oracle.py, here in torch so that autograd carries the loss into the critic's parameters).

Pinning: the dense loss part is checked against oracle/mmd_oracle.py (itself pinned to reference-minted golden
fixtures) in tests/test_scaling_cpu.py; the scaling arithmetic has no reference-side fixture (the reference needs a
TF session + conv critic to evaluate it): parity unpinned for model.py:366-403 beyond this line-by-line restatement
and a finite-difference check of the Jacobian norm.  Only tests/ import this file.
"""
from __future__ import annotations

import torch


def squared_norm_jacobian(y, x):
    d = y.shape[1]
    red = tuple(range(1, x.dim()))
    norm_gradients = torch.stack(
        [(torch.autograd.grad(y[:, i].sum(), x, create_graph=True, retain_graph=True)[0] ** 2).sum(dim=red) for i in range(d)])
    return norm_gradients.sum(dim=0)


def scale_of(x_hat, x_hat_data, sc=10.0, variant="grad"):
    norm2_jac = squared_norm_jacobian(x_hat, x_hat_data).mean()
    norm_discriminator = (x_hat ** 2).mean()
    if variant == "grad":
        return 1.0 / (sc * norm2_jac + 1.0), norm2_jac, norm_discriminator
    return 1.0 / (sc * (norm2_jac + norm_discriminator) + 1.0), norm2_jac, norm_discriminator


def _blocks(X, Y):
    XX, XY, YY = X @ X.T, X @ Y.T, Y @ Y.T
    return XX, XY, YY, torch.diagonal(XX), torch.diagonal(YY)


def mmd2_dense(kernel, X, Y, biased=False, sigma=1.0, alphas=(0.1, 1.0, 10.0)):
    """mmd2(_rbf_kernel(X, Y)) or mmd2(_mix_rq_kernel(X, Y)) on torch tensors (differentiable)."""
    XX, XY, YY, nx, ny = _blocks(X, Y)

    def k(G, nr, nc):
        D = torch.clamp(-2.0 * G + nr[:, None] + nc[None, :], min=0.0)
        if kernel == "rbf":
            return torch.exp(-D / (2.0 * sigma ** 2))
        return sum(torch.exp(-a * torch.log(1.0 + D / (2.0 * a))) for a in alphas)

    K_XX, K_XY, K_YY = k(XX, nx, nx), k(XY, nx, ny), k(YY, ny, ny)
    cd = 1.0 if kernel == "rbf" else float(len(alphas))
    m, n = float(X.shape[0]), float(Y.shape[0])
    if biased:
        return K_XX.sum() / (m * m) + K_YY.sum() / (n * n) - 2 * K_XY.sum() / (m * n)
    return ((K_XX.sum() - m * cd) / (m * (m - 1)) + (K_YY.sum() - n * cd) / (n * (n - 1)) - 2 * K_XY.sum() / (m * n))


def smmd_losses(critic, fake, images, kernel="rbf", sc=10.0, variant="grad"):
    """SMMD.set_loss + add_scaling: returns (g_loss, d_loss, scale, unscaled mmd2)."""
    images = images.requires_grad_(True)
    d_images = critic(images)
    d_G = critic(fake)
    mmd2 = mmd2_dense(kernel, d_G, d_images)
    scale, _, _ = scale_of(d_images, images, sc, variant)
    g_loss = mmd2 * scale
    return g_loss, -g_loss, scale, mmd2
