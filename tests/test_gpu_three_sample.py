"""GPU: smmd_poly_sums / the 3-sample test API (reference gan/core/mmd.py:296-539) against the oracle
and the reference's golden outputs, through the C ABI."""
import numpy as np
import pytest
import torch

from golden_util import load_three_sample_golden, three_sample_codes
from oracle import three_sample_oracle as tso

pytestmark = pytest.mark.gpu
Z3 = load_three_sample_golden()


def _check_sums(got, want, rtol):
    for g, w in zip(got, want):
        g = np.asarray(g, dtype=np.float64)
        w = np.asarray(w, dtype=np.float64)
        assert np.abs(g - w).max() <= rtol * np.abs(w).max(), (np.abs(g - w).max(), np.abs(w).max())


@pytest.mark.parametrize("tag", ["small", "mid"])
@pytest.mark.parametrize("precision,rtol", [("fp32", 1e-5), ("bf16x3", 1e-4)])
def test_sums_and_statistic_vs_reference_golden(tag, precision, rtol):
    """numpy in -> numpy out, same call sequence as scorer.py:129,157.  Tolerances: the exact path reproduces the
    reference's fp32 sums to 1e-5; the split-bf16 Gram to 1e-4 (cubing triples the 2^-17 input rounding).
    The statistic divides by sqrt(var) where var cancels heavily -> looser."""
    from smmd import _lib, mmd

    X, Y, Z = three_sample_codes(Z3, tag, np.float32)
    zs = mmd.np_diff_polynomial_mmd2_and_ratio_with_saving(X, Z, None, precision=precision)
    assert _lib.last_path() in ("simt_fp32_kid", "tc_bf16x3_kid")
    diff, ratio, ys = mmd.np_diff_polynomial_mmd2_and_ratio_with_saving(X, Y, zs, precision=precision)
    assert isinstance(diff, float) and isinstance(ys[0], np.ndarray)
    for name, sums in (("ys", ys), ("zs", zs)):
        vec = Z3["%s_f64_%s_vec" % (name, tag)]
        sc = Z3["%s_f64_%s_sc" % (name, tag)]
        _check_sums((sums[0], sums[2], sums[3], sums[1], sums[4]), (vec[0], vec[1], vec[2], sc[0], sc[1]), rtol)
    ref = Z3["res_f64_%s" % tag]
    assert abs(diff - ref[0]) <= 50 * rtol * abs(ref[0]), (diff, ref[0])
    assert abs(ratio - ref[1]) <= 500 * rtol * abs(ref[1]), (ratio, ref[1])


def test_scorer_size_vs_oracle():
    """2048 x 2048 codes (scorer.py:121 bs = 2048): default precision, torch API, against the fp64 oracle."""
    from smmd import mmd

    rng = np.random.RandomState(7)
    X = np.maximum(rng.randn(2048, 2048), 0).astype(np.float32)
    Y = np.maximum(rng.randn(2048, 2048) + 0.03, 0).astype(np.float32)
    Z = np.maximum(rng.randn(2048, 2048) + 0.06, 0).astype(np.float32)
    Xt, Yt, Zt = (torch.tensor(a, device="cuda") for a in (X, Y, Z))
    diff, ratio = mmd.diff_polynomial_mmd2_and_ratio(Xt, Yt, Zt)
    X64, Y64, Z64 = X.astype(np.float64), Y.astype(np.float64), Z.astype(np.float64)
    ys = tso.related_sums(tso.cubic_kernel(X64, Y64), tso.cubic_kernel(Y64, Y64))
    zs = tso.related_sums(tso.cubic_kernel(X64, Z64), tso.cubic_kernel(Z64, Z64))
    want = tso.diff_from_sums(ys, zs, 2048.0, graph=True)   # torch tensors in: the graph variant's clamp (mmd.py:398)
    got_ys = mmd.polynomial_related_sums(Xt, Yt)
    _check_sums([g.cpu().numpy() for g in got_ys], ys, 1e-4)
    assert abs(diff.item() - want[0]) <= 5e-3 * abs(want[0]), (diff.item(), want[0])
    assert abs(ratio.item() - want[1]) <= 5e-2 * abs(want[1]), (ratio.item(), want[1])
    # saved-sums form gives the same numbers as the three-matrix form
    zs_gpu = mmd.polynomial_related_sums(Xt, Zt)
    d2, r2, _ = mmd.diff_polynomial_mmd2_and_ratio_with_saving(Xt, Yt, zs_gpu)
    assert d2.item() == diff.item() and r2.item() == ratio.item()


def test_properties_and_errors():
    from smmd import _lib, mmd

    rng = np.random.RandomState(3)
    A = torch.tensor(np.maximum(rng.randn(300, 100), 0).astype(np.float32), device="cuda")
    B = torch.tensor(np.maximum(rng.randn(300, 100) + 0.1, 0).astype(np.float32), device="cuda")
    s = mmd.polynomial_related_sums(A, B, precision="fp32")
    # sum over rows of K_XY.sum(1) == sum over columns of K_XY.sum(0)
    assert abs(s[2].sum().item() - s[3].sum().item()) <= 1e-9 * abs(s[2].sum().item())
    # Y vs itself: difference 0, statistic 0
    d, r = mmd.diff_polynomial_mmd2_and_ratio(A, B, B, precision="fp32")
    assert d.item() == 0.0 and r.item() == 0.0
    # row permutation of Y permutes the per-row vectors and leaves the scalars (up to fp64 summation order)
    perm = torch.randperm(300, device="cuda")
    sp = mmd.polynomial_related_sums(A, B[perm], precision="fp32")
    assert (sp[0] - s[0][perm]).abs().max() <= 1e-9 * s[0].abs().max()
    assert (sp[2] - s[2][perm]).abs().max() <= 1e-9 * s[2].abs().max()
    assert abs(sp[1].item() - s[1].item()) <= 1e-9 * abs(s[1].item())
    with pytest.raises(ValueError):
        mmd.polynomial_related_sums(A, B[:200])
    with pytest.raises(RuntimeError):
        mmd.polynomial_related_sums(A.cpu(), B.cpu())
