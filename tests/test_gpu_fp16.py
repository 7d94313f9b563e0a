"""GPU parity of the fp16 operand tier (`precision="fp16"`, SMMD_PREC_FP16): the tensor-core gradient paths with IEEE-half
operands (features and the weight matrix W) against the fp64 numpy oracle, through the Python drop-in -> C ABI.

This is the tier that meets north_star's 1e-3 for GRADIENTS as well: every element |g - g64| <= 1e-3 * max|g64| (measured
1.3e-4 .. 3.3e-4; the bf16 tier sits at 1e-3 .. 2e-3) and ||g - g64||_F <= 5e-4 ||g64||_F; MMD^2
within 1e-3 relative (measured ~2e-5).  All four kernels families of paths are covered: fused symmetric (d <= 256),
symmetric two-pass (d > 256), and the row-stacked fused / two-pass kernels that row shards use."""
import numpy as np
import pytest
import torch

from oracle import mmd_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _data(m, n, d, seed):
    rng = np.random.RandomState(seed)
    X = (rng.randn(m, d) / np.sqrt(d)).astype(np.float32)
    Y = ((1.05 * rng.randn(n, d) + 0.1) / np.sqrt(d)).astype(np.float32)
    return X, Y


def _kscale(name, kw, X, Y):
    n = min(len(X), 256)
    Kxx, Kxy, Kyy, _ = mmd_oracle.kernel_matrices(name, X[:n], Y[:n], np.float64, **kw)
    return max(abs(Kxx).mean(), abs(Kxy).mean(), abs(Kyy).mean())


PATHS = {   # name -> (options, expected path for d <= 256, for d > 256)
    "symmetric": ({"sym": 1, "sym_min_rows": 1, "symf_min_rows": 1}, "tc_bf16_symf", "tc_bf16_sym"),
    "row_stacked": ({"sym": 0}, "tc_bf16_fused", "tc_bf16_wz"),
}
SHAPES = [(300, 200, 100), (1000, 1100, 256), (513, 700, 192), (700, 900, 512), (640, 520, 1024), (2048, 2048, 64)]
CASES = [("mix_rq", {}), ("rbf", {}), ("mix_rbf", {"sigmas": [1.0, 2.0, 4.0, 8.0, 16.0]}), ("mix_rq_dot", {}),
         ("tanh_mix_rq", {}), ("distance", {}), ("mix_rq", {"alphas": [0.2, 0.5, 1.0, 2.0, 5.0], "wts": [1.0, 0.5, 2.0, 1.0, 0.25]})]


@pytest.mark.parametrize("path", sorted(PATHS))
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("case", CASES, ids=lambda c: c[0] + ("+" if c[1] else ""))
def test_fp16_tier_vs_oracle(case, shape, path):
    from smmd import _lib, mmd

    name, kw = case
    m, n, d = shape
    opts, want_narrow, want_wide = PATHS[path]
    X, Y = _data(m, n, d, m + n + d)
    for k, v in opts.items():
        _lib.set_option(k, v)
    try:
        for biased in (False, True):
            Xt = torch.tensor(X, device=DEV, requires_grad=True)
            Yt = torch.tensor(Y, device=DEV, requires_grad=True)
            loss = mmd.mmd2(getattr(mmd, "_%s_kernel" % name)(Xt, Yt, **kw), biased=biased, precision="fp16")
            loss.backward()
            assert _lib.last_path() == (want_narrow if d <= 256 else want_wide)
            v, gx, gy = mmd_oracle.mmd2_and_grads(name, X, Y, biased, np.float64, **kw)
            assert abs(loss.item() - v) <= 1e-3 * abs(v) + 2e-6 * _kscale(name, kw, X, Y), (name, biased, loss.item(), v)
            for got, ref in ((Xt.grad, gx), (Yt.grad, gy)):
                diff = got.cpu().numpy().astype(np.float64) - ref
                assert np.abs(diff).max() <= 1e-3 * np.abs(ref).max(), (name, biased, np.abs(diff).max(), np.abs(ref).max())
                assert np.linalg.norm(diff) <= 5e-4 * np.linalg.norm(ref), (name, biased)
    finally:
        _lib.set_option("sym", 1)
        _lib.set_option("sym_min_rows", 0)
        _lib.set_option("symf_min_rows", 0)


def test_fp16_tier_is_explicit_and_gradient_only():
    """AUTO never picks fp16; a forward-only fp16 request is refused (never silently served by another tier)."""
    from smmd import _lib, mmd

    X, Y = _data(600, 500, 64, 3)
    Xt, Yt = torch.tensor(X, device=DEV), torch.tensor(Y, device=DEV)
    spec = mmd._mix_rq_kernel(Xt, Yt).spec
    mmd.fused_mmd2_raw(spec, Xt, Yt)                    # precision auto
    assert _lib.last_path() in ("tc_bf16_fused", "tc_bf16_sym", "tc_bf16_symf")
    with pytest.raises(_lib.SmmdError):
        mmd.fused_mmd2_raw(spec, Xt, Yt, want_grad=False, precision="fp16")


def test_fp16_range_overflow_is_reported_not_hidden():
    """IEEE half tops out at 65504: a feature beyond it becomes inf in the operand matrix and the call reports it through
    SMMD_S_NONFINITE (the bf16 tier has fp32's exponent range and accepts the same input)."""
    from smmd import _lib, mmd

    X, Y = _data(600, 500, 64, 4)
    X[7, 3] = 1.0e5
    Xt, Yt = torch.tensor(X, device=DEV), torch.tensor(Y, device=DEV)
    spec = mmd._rbf_kernel(Xt, Yt).spec
    sc, _, _ = mmd.fused_mmd2_raw(spec, Xt, Yt, precision="fp16")
    assert sc[_lib.S_NONFINITE].item() != 0.0
    sc, _, _ = mmd.fused_mmd2_raw(spec, Xt, Yt, precision="bf16")
    assert sc[_lib.S_NONFINITE].item() == 0.0


def test_fp16_large_batch_scaling_of_w():
    """At large N the weights W ~ 1/N^2 are far below half's smallest normal: they are carried times a power of two
    (w_scale_for) and the gradients must still match the exact path, here at 16384 + 16384 rows (W ~ 4e-9)."""
    from smmd import _lib, mmd

    g = torch.Generator(device=DEV).manual_seed(11)
    n, d = 16384, 128
    Xt = torch.randn(n, d, device=DEV, generator=g) / d ** 0.5
    Yt = (1.05 * torch.randn(n, d, device=DEV, generator=g) + 0.1) / d ** 0.5
    spec = mmd._mix_rq_kernel(Xt, Yt).spec
    a, gX, gY = mmd.fused_mmd2_raw(spec, Xt, Yt, precision="fp16")
    assert _lib.last_path() == "tc_bf16_sym"
    b, rX, rY = mmd.fused_mmd2_raw(spec, Xt, Yt, precision="fp32")
    assert abs(a[_lib.S_MMD2].item() - b[_lib.S_MMD2].item()) <= 1e-3 * abs(b[_lib.S_MMD2].item())
    for got, ref in ((gX, rX), (gY, rY)):
        assert (got - ref).abs().max() <= 1e-3 * ref.abs().max()
