"""CPU: the 3-sample oracle against the reference's own outputs (golden fixtures), and the product's host-side
estimator arithmetic (smmd/mmd.py, no GPU needed for that part) against the oracle."""
import numpy as np
import pytest

from golden_util import load_three_sample_golden, three_sample_codes
from oracle import three_sample_oracle as tso

Z3 = load_three_sample_golden()


@pytest.mark.parametrize("tag", ["small", "mid"])
@pytest.mark.parametrize("dn,dt,rtol", [("f32", np.float32, 1e-5), ("f64", np.float64, 1e-12)])
def test_oracle_matches_reference(tag, dn, dt, rtol):
    X, Y, Z = three_sample_codes(Z3, tag, dt)
    zs = tso.diff_with_saving(X, Z, None)
    diff, ratio, ys = tso.diff_with_saving(X, Y, zs)
    ref = Z3["res_%s_%s" % (dn, tag)]
    assert abs(diff - ref[0]) <= rtol * abs(ref[0])
    assert abs(ratio - ref[1]) <= rtol * abs(ref[1])
    for name, sums in (("ys", ys), ("zs", zs)):
        vec = Z3["%s_%s_%s_vec" % (name, dn, tag)]
        sc = Z3["%s_%s_%s_sc" % (name, dn, tag)]
        got = np.stack([sums[0], sums[2], sums[3]]).astype(np.float64)
        assert np.abs(got - vec).max() <= rtol * np.abs(vec).max()
        assert abs(sums[1] - sc[0]) <= rtol * abs(sc[0]) and abs(sums[4] - sc[1]) <= rtol * abs(sc[1])


def test_host_estimator_arithmetic_matches_oracle():
    from smmd import mmd

    X, Y, Z = three_sample_codes(Z3, "mid", np.float64)
    ys = tso.related_sums(tso.cubic_kernel(X, Y), tso.cubic_kernel(Y, Y))
    zs = tso.related_sums(tso.cubic_kernel(X, Z), tso.cubic_kernel(Z, Z))
    want = tso.diff_from_sums(ys, zs, float(len(Y)))
    got = mmd._np_diff_mmd2_and_ratio_from_sums(ys, zs, float(len(Y)))
    assert abs(got[0] - want[0]) <= 1e-12 * abs(want[0]) and abs(got[1] - want[1]) <= 1e-10 * abs(want[1])
    # dense-block helper (mmd.py:515-539), including the const_diagonal branch
    for cd in (False, 1.5):
        a = mmd._np_get_sums(tso.cubic_kernel(X, Y), tso.cubic_kernel(Y, Y), cd)
        b = tso.related_sums(tso.cubic_kernel(X, Y), tso.cubic_kernel(Y, Y), cd)
        for u, v in zip(a, b):
            assert np.allclose(u, v, rtol=1e-12, atol=0)


def test_variance_floor():
    """ratio = diff / sqrt(max(var, 1e-5)) (mmd.py:401/511): identical samples -> diff 0, ratio 0, no NaN."""
    X, Y, _ = three_sample_codes(Z3, "small", np.float64)
    ys = tso.related_sums(tso.cubic_kernel(X, Y), tso.cubic_kernel(Y, Y))
    diff, ratio = tso.diff_from_sums(ys, ys, float(len(Y)))
    assert diff == 0.0 and ratio == 0.0


def test_dense_block_entry_point_matches_oracle():
    """_diff_mmd2_and_ratio(K_XY, K_XZ, K_YY, K_ZZ) (mmd.py:322-336) on numpy blocks."""
    from smmd import mmd

    X, Y, Z = three_sample_codes(Z3, "small", np.float64)
    kxy, kxz = tso.cubic_kernel(X, Y), tso.cubic_kernel(X, Z)
    kyy, kzz = tso.cubic_kernel(Y, Y), tso.cubic_kernel(Z, Z)
    got = mmd._diff_mmd2_and_ratio(kxy, kxz, kyy, kzz)
    want = tso.diff_from_sums(tso.related_sums(kxy, kyy), tso.related_sums(kxz, kzz), float(len(Y)))
    assert abs(got[0] - want[0]) <= 1e-12 * abs(want[0]) and abs(got[1] - want[1]) <= 1e-10 * abs(want[1])
    ref = Z3["res_f64_small"]
    assert abs(got[0] - ref[0]) <= 1e-10 * abs(ref[0]) and abs(got[1] - ref[1]) <= 1e-8 * abs(ref[1])


@pytest.mark.parametrize("tag", ["small", "mid"])
def test_graph_variant_ratio_clamp(tag):
    """The TF variant (mmd.py:339-399) divides by mysqrt(max(var, eps)) = sqrt(max(var, eps) + eps); the fixture holds
    the reference's own output (executed over the torch-backed tf shim).  Oracle and the product's tensor branch."""
    import torch

    from smmd import mmd

    X, Y, Z = three_sample_codes(Z3, tag, np.float64)
    ys = tso.related_sums(tso.cubic_kernel(X, Y), tso.cubic_kernel(Y, Y))
    zs = tso.related_sums(tso.cubic_kernel(X, Z), tso.cubic_kernel(Z, Z))
    ref = Z3["res_graph_f64_%s" % tag]
    want = tso.diff_from_sums(ys, zs, float(len(Y)), graph=True)
    assert abs(want[0] - ref[0]) <= 1e-12 * abs(ref[0]) and abs(want[1] - ref[1]) <= 1e-10 * abs(ref[1])
    tys = tuple(torch.as_tensor(v, dtype=torch.float64) for v in ys)
    tzs = tuple(torch.as_tensor(v, dtype=torch.float64) for v in zs)
    got = mmd._diff_mmd2_and_ratio_from_sums(tys, tzs, float(len(Y)))
    assert abs(float(got[0]) - ref[0]) <= 1e-12 * abs(ref[0]) and abs(float(got[1]) - ref[1]) <= 1e-10 * abs(ref[1])
