"""GPU: K_XY_only witness block with first and second derivatives (gradient penalty, gan/core/model.py:327-350)
against torch-fp64 autograd through the reference-pinned witness oracle."""
import numpy as np
import pytest
import torch

from oracle import witness_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"
KERNELS = [("rbf", {}), ("mix_rbf", {}), ("mix_rbf", {"sigmas": [1.0, 2.0, 4.0], "wts": [1.0, 0.5, 2.0]}), ("mix_rq", {}),
           ("mix_rq_dot", {}), ("mix_rq_1dot", {}), ("dot", {}), ("distance", {}), ("tanh_distance", {}),
           ("tanh_mix_rq", {})]


def _data(m, n, d, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    X = torch.randn(m, d, generator=g, dtype=torch.float64) * scale
    Y = (1.1 * torch.randn(n, d, generator=g, dtype=torch.float64) + 0.1) * scale
    return X, Y


@pytest.mark.parametrize("shape", [(64, 64, 16), (37, 50, 7), (64, 48, 128)], ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("case", KERNELS, ids=lambda c: c[0] + ("+" if c[1] else ""))
def test_first_and_second_order_vjp(case, shape):
    """fp32 tolerance: 2e-5 of the largest reference entry for values / first derivatives, 1e-4 for the second
    derivatives (products of two fp32 pair sums)."""
    from smmd import mmd

    name, kw = case
    m, n, d = shape
    X64, Y64 = _data(m, n, d, m + n + d, scale=1.0 / np.sqrt(d) if name != "dot" else 1.0)
    g = torch.Generator().manual_seed(99)
    dK64 = torch.randn(m, n, generator=g, dtype=torch.float64) / n
    VX64 = torch.randn(m, d, generator=g, dtype=torch.float64)
    VY64 = torch.randn(n, d, generator=g, dtype=torch.float64)

    def run(kernel_fn, X, Y, dK, VX, VY):
        X = X.clone().requires_grad_(True)
        Y = Y.clone().requires_grad_(True)
        dK = dK.clone().requires_grad_(True)
        K = kernel_fn(X, Y)
        gX, gY = torch.autograd.grad((K * dK).sum(), [X, Y], create_graph=True)
        L = (gX * VX).sum() + (gY * VY).sum()
        hX, hY, hK = torch.autograd.grad(L, [X, Y, dK])
        return K.detach(), gX.detach(), gY.detach(), hX, hY, hK

    ref = run(lambda a, b: witness_oracle.kernel_xy(name, a, b, **kw), X64, Y64, dK64, VX64, VY64)
    f32 = lambda t: t.to(DEV, torch.float32)
    got = run(lambda a, b: getattr(mmd, "_%s_kernel" % name)(a, b, K_XY_only=True, **kw), f32(X64), f32(Y64), f32(dK64),
              f32(VX64), f32(VY64))
    tols = [2e-5, 2e-5, 2e-5, 1e-4, 1e-4, 1e-4]
    for what, a, b, tol in zip(("K", "dX", "dY", "d2X", "d2Y", "d2dK"), got, ref, tols):
        err = (a.double().cpu() - b).abs().max().item()
        assert err <= tol * max(b.abs().max().item(), 1e-6), (name, what, err, b.abs().max().item())


def test_gradient_penalty_like_the_reference():
    """model.py:327-350 with a small torch critic: witness on interpolates, gradient w.r.t. the critic INPUT,
    penalty = mean((|grad| - 1)^2), then the gradient of the penalty w.r.t. the critic WEIGHTS (double backward
    through the witness kernels) -- product path vs fp64 oracle."""
    from smmd import mmd

    torch.manual_seed(5)
    bs, din, dof = 64, 48, 16
    W1 = torch.randn(din, 32, dtype=torch.float64) / np.sqrt(din)
    W2 = torch.randn(32, dof, dtype=torch.float64) / np.sqrt(32)
    real_data = torch.rand(bs, din, dtype=torch.float64)
    fake_data = torch.rand(bs, din, dtype=torch.float64) * 0.8
    alpha = torch.rand(bs, 1, dtype=torch.float64)

    def penalty(kernel_xy, W1, W2, real_data, fake_data, alpha):
        critic = lambda z: torch.nn.functional.softplus(z @ W1) @ W2
        x_hat_data = ((1.0 - alpha) * real_data + alpha * fake_data).requires_grad_(True)
        x_hat, real, fake = critic(x_hat_data), critic(real_data), critic(fake_data)
        wit = kernel_xy(x_hat, real).mean(1) - kernel_xy(x_hat, fake).mean(1)
        grads = torch.autograd.grad(wit.sum(), x_hat_data, create_graph=True)[0]
        norm = torch.sqrt((grads * grads).sum(1) + 1e-12)
        return ((norm - 1.0) ** 2).mean()

    for name in ("mix_rq", "rbf", "distance"):
        Wr = [W1.clone().requires_grad_(True), W2.clone().requires_grad_(True)]
        p_ref = penalty(lambda a, b: witness_oracle.kernel_xy(name, a, b), Wr[0], Wr[1], real_data, fake_data, alpha)
        g_ref = torch.autograd.grad(p_ref, Wr)
        c = lambda t: t.to(DEV, torch.float32)
        Wg = [c(W1).requires_grad_(True), c(W2).requires_grad_(True)]
        p_got = penalty(lambda a, b: getattr(mmd, "_%s_kernel" % name)(a, b, K_XY_only=True), Wg[0], Wg[1], c(real_data),
                        c(fake_data), c(alpha))
        g_got = torch.autograd.grad(p_got, Wg)
        assert abs(p_got.item() - p_ref.item()) <= 1e-4 * abs(p_ref.item()), (name, p_got.item(), p_ref.item())
        for a, b in zip(g_got, g_ref):
            err = (a.double().cpu() - b).abs().max().item()
            assert err <= 2e-3 * b.abs().max().item(), (name, err, b.abs().max().item())
