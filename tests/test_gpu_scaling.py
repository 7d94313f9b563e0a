"""GPU: the SMMD scaled loss (smmd.scaling.scaled_mmd2 over the fused MMD^2 op) through a small conv critic, against
oracle/scaling_oracle.py (torch fp64 on the CPU, dense kernels + autograd, same weights): loss value, scale, and the
gradient of d_loss w.r.t. every critic parameter -- i.e. scale * dMMD2 + mmd2 * dscale arrives intact
(gan/core/smmd.py:10-23, gan/core/model.py:366-403)."""
import copy

import numpy as np
import pytest
import torch

from oracle import scaling_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _critic(dof):
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.LeakyReLU(0.2), torch.nn.Conv2d(8, 8, 4, 2, 1),
                               torch.nn.LeakyReLU(0.2), torch.nn.Flatten(), torch.nn.Linear(8 * 8 * 8, dof))


@pytest.mark.parametrize("kernel,dof", [("rbf", 1), ("mix_rq", 16)])
@pytest.mark.parametrize("variant", ["grad", "value_and_grad"])
def test_scaled_loss_parameter_gradients_vs_oracle(kernel, dof, variant):
    from smmd import mmd, scaling

    # the critic is plain PyTorch/cuDNN: keep its convolutions in true fp32 (TF32 is on by default for cuDNN convs and
    # would put 1e-3 differences into the features before the loss under test ever sees them)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    net = _critic(dof)
    ref_net = copy.deepcopy(net).double()
    net = net.to(DEV)
    rs = np.random.RandomState(5)
    images = rs.rand(64, 3, 16, 16).astype(np.float32)
    fake = rs.rand(64, 3, 16, 16).astype(np.float32) * 0.9 + 0.05
    # ---- product path ----
    x = torch.tensor(images, device=DEV, requires_grad=True)
    d_images = net(x)
    d_G = net(torch.tensor(fake, device=DEV))
    scale, nj, nd = scaling.smmd_scale(d_images, x, scaling_coeff=10.0, scaling_variant=variant)
    g_loss, unscaled = scaling.scaled_mmd2(getattr(mmd, "_%s_kernel" % kernel)(d_G, d_images), scale, precision="fp32")
    (-g_loss).backward()
    # ---- oracle ----
    rg, rd, rscale, rmmd2 = scaling_oracle.smmd_losses(ref_net, torch.tensor(fake, dtype=torch.float64),
                                                       torch.tensor(images, dtype=torch.float64), kernel, 10.0, variant)
    rd.backward()
    assert abs(float(scale) - float(rscale)) <= 1e-4 * float(rscale)
    assert abs(float(unscaled) - float(rmmd2)) <= 1e-5 * abs(float(rmmd2)) + 1e-7
    assert abs(float(g_loss) - float(rg)) <= 1e-4 * abs(float(rg)) + 1e-7
    # (the last bias has an exactly zero gradient for distance-based kernels -- a common shift of all features changes
    # nothing -- so every parameter is compared on the scale of the largest gradient of its kind, not its own)
    gmax = max(float(q.grad.abs().max()) for q in ref_net.parameters())
    for (name, p), q in zip(net.named_parameters(), ref_net.parameters()):
        got, ref = p.grad.double().cpu(), q.grad
        tol = 2e-4 * float(ref.abs().max()) + 2e-5 * gmax   # (second term: fp32 noise of the critic's backward)
        assert (got - ref).abs().max() <= tol, (name, float((got - ref).abs().max()), float(ref.abs().max()), gmax)


def test_scaled_mmd2_backward_is_scale_times_dx_and_mmd2():
    """The single backward node returns exactly (scale * dX, scale * dY, mmd2)."""
    from smmd import _lib, mmd, scaling

    g = torch.Generator(device=DEV).manual_seed(1)
    X = (torch.randn(2000, 64, device=DEV, generator=g) / 8).requires_grad_(True)
    Y = ((1.05 * torch.randn(2100, 64, device=DEV, generator=g) + 0.1) / 8).requires_grad_(True)
    scale = torch.tensor(0.37, device=DEV, requires_grad=True)
    loss, unscaled = scaling.scaled_mmd2(mmd._mix_rq_kernel(X, Y), scale, precision="bf16")
    loss.backward()
    sc, dX, dY = mmd.fused_mmd2_raw(mmd._mix_rq_kernel(X, Y).spec, X, Y, precision="bf16")
    assert _lib.last_path() in ("tc_bf16_fused", "tc_bf16_sym")   # 4224 stacked rows: the symmetric path by default
    assert torch.allclose(X.grad, 0.37 * dX, rtol=1e-6, atol=0) and torch.allclose(Y.grad, 0.37 * dY, rtol=1e-6, atol=0)
    assert abs(float(scale.grad) - float(sc[_lib.S_MMD2])) <= 1e-6 * abs(float(sc[_lib.S_MMD2]))
    assert abs(float(loss) - 0.37 * float(unscaled)) <= 1e-6 * abs(float(loss))
