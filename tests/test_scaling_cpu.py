"""CPU: smmd.scaling (SMMD gradient-norm scaling, host-side autograd code) against oracle/scaling_oracle.py, the
oracle's dense loss against oracle/mmd_oracle.py, and the Jacobian norm against finite differences."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

from oracle import mmd_oracle, scaling_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _scaling_module():
    """smmd.scaling imports the ctypes binding lazily-safe modules only; load it without a GPU."""
    sys.path.insert(0, os.path.join(ROOT, "scaled-mmd-gan_b200"))
    from smmd import scaling
    return scaling


def _critic(dof, seed=0, conv=False):
    torch.manual_seed(seed)
    if conv:
        net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1), torch.nn.LeakyReLU(0.2), torch.nn.Conv2d(4, 4, 4, 2, 1),
                                  torch.nn.LeakyReLU(0.2), torch.nn.Flatten(), torch.nn.Linear(4 * 4 * 4, dof))
    else:
        net = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(3 * 8 * 8, 16), torch.nn.Tanh(), torch.nn.Linear(16, dof))
    return net.double()


@pytest.mark.parametrize("dof", [1, 3, 16])
@pytest.mark.parametrize("conv", [False, True])
def test_squared_norm_jacobian_matches_reference_loop(dof, conv):
    scaling = _scaling_module()
    net = _critic(dof, conv=conv)
    x = torch.randn(5, 3, 8, 8, dtype=torch.float64, requires_grad=True)
    got = scaling.squared_norm_jacobian(net(x), x)
    ref = scaling_oracle.squared_norm_jacobian(net(x), x)
    assert torch.allclose(got, ref, rtol=1e-12, atol=1e-14)
    # and it stays differentiable w.r.t. the critic parameters (the scale is trained through)
    g1 = torch.autograd.grad(got.mean(), list(net.parameters()), allow_unused=True)
    g2 = torch.autograd.grad(ref.mean(), list(net.parameters()), allow_unused=True)
    for a, b in zip(g1, g2):
        if a is None or b is None:   # the last bias does not enter the Jacobian
            assert a is None and b is None or (a is None and float(b.abs().max()) == 0) or (b is None and float(a.abs().max()) == 0)
        else:
            assert torch.allclose(a, b, rtol=1e-10, atol=1e-13)


def test_squared_norm_jacobian_finite_differences():
    scaling = _scaling_module()
    net = _critic(2)
    x = torch.randn(3, 3, 8, 8, dtype=torch.float64, requires_grad=True)
    got = scaling.squared_norm_jacobian(net(x), x).detach()
    eps = 1e-6
    fd = torch.zeros(3, dtype=torch.float64)
    xf = x.detach().reshape(3, -1)
    for j in range(xf.shape[1]):
        e = torch.zeros_like(xf)
        e[:, j] = eps
        dy = (net((xf + e).reshape(x.shape)) - net((xf - e).reshape(x.shape))) / (2 * eps)
        fd += (dy ** 2).sum(dim=1).detach()
    assert torch.allclose(got, fd, rtol=1e-6)


@pytest.mark.parametrize("variant", ["grad", "value_and_grad"])
def test_scale_matches_oracle(variant):
    scaling = _scaling_module()
    net = _critic(4, conv=True)
    x = torch.rand(6, 3, 8, 8, dtype=torch.float64, requires_grad=True)
    s, nj, nd = scaling.smmd_scale(net(x), x, scaling_coeff=10.0, scaling_variant=variant)
    rs, rnj, rnd = scaling_oracle.scale_of(net(x), x, 10.0, variant)
    assert abs(float(s) - float(rs)) <= 1e-13 and abs(float(nj) - float(rnj)) <= 1e-12 * float(rnj)
    with pytest.raises(ValueError):
        scaling.smmd_scale(net(x), x, scaling_variant="nope")


@pytest.mark.parametrize("kernel,name", [("rbf", "rbf"), ("mix_rq", "mix_rq")])
def test_oracle_dense_loss_matches_mmd_oracle(kernel, name):
    rs = np.random.RandomState(3)
    X = rs.randn(40, 6) / 2.5
    Y = (1.05 * rs.randn(36, 6) + 0.1) / 2.5
    for biased in (False, True):
        Xt = torch.tensor(X, requires_grad=True)
        Yt = torch.tensor(Y, requires_grad=True)
        v = scaling_oracle.mmd2_dense(kernel, Xt, Yt, biased=biased)
        gx, gy = torch.autograd.grad(v, [Xt, Yt])
        rv, rgx, rgy = mmd_oracle.mmd2_and_grads(name, X, Y, biased, np.float64)
        assert abs(float(v) - rv) <= 1e-12 * max(1.0, abs(rv))
        assert np.abs(gx.numpy() - rgx).max() <= 1e-12 and np.abs(gy.numpy() - rgy).max() <= 1e-12
