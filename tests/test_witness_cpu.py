"""CPU: the differentiable witness oracle reproduces the reference's K_XY_only outputs (golden fixtures)."""
import numpy as np
import torch

from golden_util import load_mmd_golden
from oracle import witness_oracle

Z, INDEX = load_mmd_golden()


def test_witness_oracle_matches_reference_kxy():
    n = 0
    for k in Z.files:
        if not k.startswith("kxy|"):
            continue
        _, shape, kname = k.split("|")
        X, Y = torch.tensor(Z["X_" + shape], dtype=torch.float64), torch.tensor(Z["Y_" + shape], dtype=torch.float64)
        K = witness_oracle.kernel_xy(kname, X, Y).numpy()
        # the fixture is the reference's fp32 output
        assert np.allclose(K, Z[k], rtol=2e-5, atol=2e-5), k
        n += 1
    assert n >= 5


def test_second_derivative_is_symmetric():
    """Sanity of the oracle itself: d2 sum(K * dK) / dX dY contracted both ways agrees (autograd double backward)."""
    g = torch.Generator().manual_seed(0)
    X = torch.randn(7, 5, generator=g, dtype=torch.float64, requires_grad=True)
    Y = torch.randn(6, 5, generator=g, dtype=torch.float64, requires_grad=True)
    dK = torch.randn(7, 6, generator=g, dtype=torch.float64)
    VX = torch.randn(7, 5, generator=g, dtype=torch.float64)
    VY = torch.randn(6, 5, generator=g, dtype=torch.float64)
    for name in ("rbf", "mix_rq_dot", "distance"):
        K = witness_oracle.kernel_xy(name, X, Y)
        gX, gY = torch.autograd.grad((K * dK).sum(), [X, Y], create_graph=True)
        a = torch.autograd.grad((gX * VX).sum(), Y, retain_graph=True)[0]
        b = torch.autograd.grad((gY * VY).sum(), X, retain_graph=True)[0]
        # <VY, d/dY <VX, gX>> == <VX, d/dX <VY, gY>>
        assert abs((a * VY).sum().item() - (b * VX).sum().item()) <= 1e-10 * (1 + abs((a * VY).sum().item()))
