"""GPU: edge shapes across the three gradient paths (single-launch exact kernel, fused tcgen05 kernel, two-pass path):
sizes around the 128-row / 64- and 256-column tile boundaries, d around 64 / 256 / 512 / 1024, very unequal m and n.
The tensor-core result is compared with the exact fp32 path on the same device (itself pinned to the reference by
test_gpu_parity.py), bf16 tolerance: 1e-3 on MMD^2 (plus the cancellation floor), 4e-3 of max|g| on gradients."""
import numpy as np
import pytest
import torch

from oracle import mmd_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _row_stacked_kernels():
    """This module pins the row-stacked kernels (fused tile-pair kernel, W row panels + GEMM): whole problems of >= 4096
    stacked rows would otherwise take the symmetric paths, which tests/test_gpu_sym.py covers."""
    from smmd import _lib

    _lib.set_option("sym", 0)
    yield
    _lib.set_option("sym", 1)
DEV = "cuda"

SHAPES = [
    (1, 300, 40), (300, 1, 40), (2, 2, 300), (127, 129, 64), (128, 128, 65), (129, 127, 63), (255, 257, 255),
    (256, 256, 256), (257, 255, 257), (383, 130, 511), (384, 640, 512), (385, 385, 513), (64, 3000, 1023),
    (3000, 64, 1025), (1000, 1000, 2049), (513, 1, 700), (640, 1281, 320), (1279, 641, 96), (2047, 2049, 33),
]


def _data(m, n, d, seed):
    rng = np.random.RandomState(seed)
    X = (rng.randn(m, d) / np.sqrt(d)).astype(np.float32)
    Y = ((1.05 * rng.randn(n, d) + 0.1) / np.sqrt(d)).astype(np.float32)
    return torch.tensor(X, device=DEV), torch.tensor(Y, device=DEV)


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("kernel", ["mix_rq", "mix_rbf", "distance"])
def test_tensor_core_paths_on_edge_shapes(kernel, shape):
    from smmd import _lib, mmd

    m, n, d = shape
    X, Y = _data(m, n, d, m * 7 + n * 3 + d)
    kw = {"sigmas": [1.0, 2.0, 4.0, 8.0, 16.0]} if kernel == "mix_rbf" else {}
    spec = getattr(mmd, "_%s_kernel" % kernel)(X, Y, **kw).spec
    for biased in ((False, True) if min(m, n) > 1 else (True,)):   # unbiased needs m, n >= 2 (division by m - 1)
        if d <= 2048:
            ref, rX, rY = mmd.fused_mmd2_raw(spec, X, Y, biased=biased, precision="fp32")
        else:   # the exact gradient kernel stops at d = 2048 (include/smmd.h): values from it, gradients from the oracle
            with pytest.raises(_lib.SmmdError):
                mmd.fused_mmd2_raw(spec, X, Y, biased=biased, precision="fp32")
            ref, _, _ = mmd.fused_mmd2_raw(spec, X, Y, biased=biased, want_grad=False, precision="fp32")
            _, gx, gy = mmd_oracle.mmd2_and_grads(kernel, X.cpu().numpy(), Y.cpu().numpy(), biased, np.float64, **kw)
            rX, rY = torch.tensor(gx, device=DEV, dtype=torch.float32), torch.tensor(gy, device=DEV, dtype=torch.float32)
            auto, _, _ = mmd.fused_mmd2_raw(spec, X, Y, biased=biased)          # AUTO: wide rows -> tensor cores
            assert _lib.last_path().startswith("tc_bf16_wz")
        got, gX, gY = mmd.fused_mmd2_raw(spec, X, Y, biased=biased, precision="bf16")
        assert _lib.last_path() in ("tc_bf16_fused", "tc_bf16_wz", "tc_bf16_wz_pair")
        assert got[_lib.S_NONFINITE].item() == 0.0
        kscale = max(abs(ref[i].item()) / cnt for i, cnt in ((_lib.S_SUM_XX, max(m * m, 1)), (_lib.S_SUM_YY, max(n * n, 1)),
                                                              (_lib.S_SUM_XY, m * n)))
        v = ref[_lib.S_MMD2].item()
        # MMD^2 is a difference of block means: with a handful of rows per set nothing averages the bf16 rounding of
        # the individual kernel values, so the floor is that rounding times the kernel scale
        floor = (4e-6 if min(m, n) >= 16 else 4e-3) * kscale
        assert abs(got[_lib.S_MMD2].item() - v) <= 1e-3 * abs(v) + floor, (kernel, shape, biased, got[_lib.S_MMD2].item(), v)
        # a set with a handful of rows leaves single-term sums W_ij (z_i - bf16(z_j)): no averaging of the 2^-9 operand
        # rounding, so the bound is the bf16 unit roundoff itself there
        gtol = 4e-3 if min(m, n) >= 16 else 1.2e-2
        for a, b in ((gX, rX), (gY, rY)):
            err, ref_max = (a - b).abs().max().item(), b.abs().max().item()
            assert err <= gtol * ref_max + 1e-12, (kernel, shape, biased, err, ref_max)
