"""SMMD gradient-norm scaling: the reference's ``squared_norm_jacobian`` (gan/core/ops.py:228-233),
the scale of ``MMD_GAN.add_scaling`` (gan/core/model.py:366-403) and ``SMMD.apply_scaling`` (gan/core/smmd.py:21-23),
with the scaled loss as ONE autograd node over the fused MMD^2 op.

    norm2_jac  = mean_b  sum_i || d critic(x)[b, i] / d x[b] ||^2            (ops.py:228-233, model.py:382-384)
    norm_disc  = mean (critic(x)^2)                                            (model.py:385)
    scale      = 1 / (sc * norm2_jac + 1)                     variant 'grad'    (model.py:387-388)
               = 1 / (sc * (norm2_jac + norm_disc) + 1)       'value_and_grad'  (model.py:389-390)
    g_loss     = mmd2 * scale,  d_loss = -g_loss                              (smmd.py:21-23)

The critic is PyTorch/cuDNN code and stays that (its Jacobian norm needs a double backward through the conv net);
what this module fuses is the loss side: ``scaled_mmd2`` runs the library's fused forward+backward kernel once and
returns, in a single backward node,

    d g_loss / d features = scale * dMMD2/dfeatures      d g_loss / d scale = mmd2

so ``scale * dX + mmd2 * dscale`` reaches the critic parameters without the separate multiply / broadcast nodes (and
their launches) autograd would otherwise put between the loss and the feature gradients.
"""
from __future__ import annotations

import torch

from . import _lib
from .mmd import KernelHandle, fused_mmd2_raw


def squared_norm_jacobian(y, x, create_graph=True):
    """ops.py:228-233: per-sample squared Frobenius norm of d y[b, :] / d x[b] -> tensor [B].

    ``y`` = critic(x) [B, d]; ``x`` the critic input [B, ...] with requires_grad.  The reference issues one
    ``tf.gradients(y[:, i], x)`` per output feature; here all d cotangents go through ONE batched backward
    (``is_grads_batched``), falling back to the per-feature loop for ops without a batching rule.  ``create_graph``
    keeps the result differentiable w.r.t. the critic parameters (the scale is trained through)."""
    if y.dim() != 2:
        raise ValueError("critic output must be [batch, dof_dim]")
    B, d = y.shape
    red = tuple(range(1, x.dim()))
    if d == 1:   # the shipped configs (dof_dim: 1): one backward
        (g,) = torch.autograd.grad(y.sum(), x, create_graph=create_graph)
        return (g * g).sum(dim=red)
    try:
        eye = torch.eye(d, dtype=y.dtype, device=y.device)
        cot = eye[:, None, :].expand(d, B, d)                     # cotangent i selects output feature i of every sample
        (g,) = torch.autograd.grad(y, x, grad_outputs=cot, is_grads_batched=True, create_graph=create_graph)
        return (g * g).sum(dim=(0,) + tuple(r + 1 for r in red))
    except RuntimeError:
        out = 0
        for i in range(d):
            (g,) = torch.autograd.grad(y[:, i].sum(), x, create_graph=create_graph, retain_graph=True)
            out = out + (g * g).sum(dim=red)
        return out


def smmd_scale(x_hat, x_hat_data, scaling_coeff=10.0, scaling_variant="grad", create_graph=True):
    """model.py:366-390 -> (scale, norm2_jac, norm_discriminator); x_hat = critic(x_hat_data)."""
    norm2_jac = squared_norm_jacobian(x_hat, x_hat_data, create_graph=create_graph).mean()
    norm_discriminator = (x_hat * x_hat).mean()
    if scaling_variant == "grad":
        scale = 1.0 / (scaling_coeff * norm2_jac + 1.0)
    elif scaling_variant == "value_and_grad":
        scale = 1.0 / (scaling_coeff * (norm2_jac + norm_discriminator) + 1.0)
    else:
        raise ValueError("scaling_variant must be 'grad' or 'value_and_grad' (model.py:387-390)")
    return scale, norm2_jac, norm_discriminator


class _ScaledMMD2(torch.autograd.Function):
    """g_loss = mmd2(K) * scale with ONE backward node: (scale * dX, scale * dY, mmd2)."""

    @staticmethod
    def forward(ctx, X, Y, scale, spec, biased, precision):
        need = X.requires_grad or Y.requires_grad
        scalars, dX, dY = fused_mmd2_raw(spec, X, Y, biased, want_grad=need, precision=precision)
        mmd2 = scalars[_lib.S_MMD2].to(torch.float32)
        ctx.save_for_backward(dX, dY, scale.detach(), mmd2)
        ctx.in_dtypes = (X.dtype, Y.dtype, scale.dtype)
        ctx.mark_non_differentiable(mmd2)
        return mmd2 * scale.detach().to(torch.float32), mmd2

    @staticmethod
    def backward(ctx, grad_loss, _grad_mmd2):
        dX, dY, scale, mmd2 = ctx.saved_tensors
        gs = (grad_loss.to(torch.float32) * scale.to(torch.float32))
        gX = (gs * dX).to(ctx.in_dtypes[0]) if dX is not None else None
        gY = (gs * dY).to(ctx.in_dtypes[1]) if dY is not None else None
        return gX, gY, (grad_loss.to(torch.float32) * mmd2).to(ctx.in_dtypes[2]).reshape(scale.shape), None, None, None


def scaled_mmd2(K, scale, biased=False, precision=None):
    """``mmd.mmd2(K) * scale`` (smmd.py:14,21-23) -> (g_loss, unscaled mmd2); ``K`` = ``mmd._<name>_kernel(G, images)``.
    ``d_loss`` is ``-g_loss``.  ``scale`` is a scalar tensor (from ``smmd_scale``) and receives ``mmd2 * grad``."""
    if not isinstance(K, KernelHandle):
        raise TypeError("scaled_mmd2 expects the handle returned by a _<name>_kernel(G, images) call")
    scale = torch.as_tensor(scale, device=K.X.device)
    loss, unscaled = _ScaledMMD2.apply(K.X, K.Y, scale, K.spec, bool(biased), precision)
    return loss, unscaled.detach()
