"""CPU: the oracle restatement must reproduce the reference's own outputs (golden fixtures minted by
running /root/reference unmodified, see oracle/make_golden.py), and be self-consistent."""
import numpy as np
import pytest

from oracle import kid_oracle, mmd_oracle
from golden_util import kid_codes, load_kid_golden, load_mmd_golden

Z, INDEX = load_mmd_golden()


def _rel(a, b):
    return abs(float(a) - float(b)) / max(abs(float(b)), 1e-30)


@pytest.mark.parametrize("case", INDEX, ids=[c["key"] for c in INDEX])
def test_mmd2_value_and_grad_match_reference(case):
    X, Y = Z["X_" + case["shape"]], Z["Y_" + case["shape"]]
    key = case["key"]
    # fp64 oracle vs fp64 reference: formulas identical -> near machine precision
    v, gx, gy = mmd_oracle.mmd2_and_grads(case["kernel"], X, Y, case["biased"], np.float64, **case["kwargs"])
    v64 = float(Z[key + "|v64"])
    assert abs(v - v64) <= 1e-12 * max(1.0, abs(v64)) + 1e-11 * abs(v64)
    gx64, gy64 = Z[key + "|gx64"], Z[key + "|gy64"]
    tol = 1e-10 if gx64.dtype == np.float64 else 2e-7  # non-C1 golden grads are stored as fp32
    assert np.abs(gx - gx64).max() <= tol * max(np.abs(gx64).max(), 1e-30) + 1e-14
    assert np.abs(gy - gy64).max() <= tol * max(np.abs(gy64).max(), 1e-30) + 1e-14
    # fp32 oracle value vs fp32 reference value: same arithmetic up to summation order
    v32 = mmd_oracle.mmd2(case["kernel"], X, Y, case["biased"], np.float32, **case["kwargs"])
    ref32 = float(Z[key + "|v32"])
    # (different BLAS / pairwise-summation order only; K entries are O(sum of weights) <= 6)
    assert abs(float(v32) - ref32) <= 1e-5 * max(1.0, abs(ref32))


def test_ratio_matches_reference():
    n = 0
    for k in Z.files:
        if not k.startswith("ratio|"):
            continue
        _, shape, kname, biased = k.split("|")
        kw = {"sigmas": [1.0, 2.0, 4.0, 8.0, 16.0]} if kname == "mix_rbf" else {}
        X, Y = Z["X_" + shape], Z["Y_" + shape]
        v, r, var = mmd_oracle.mmd2_and_ratio(kname, X, Y, bool(int(biased)), dtype=np.float32, **kw)
        gv, gr, gvar = Z[k]
        assert abs(v - gv) <= 1e-5 * max(abs(gv), 1.0), k   # fp32 summation-order noise on O(m^2) sums
        # the variance estimate is a difference of large fp32 terms: compare on the fp64 oracle scale
        v6, r6, var6 = mmd_oracle.mmd2_and_ratio(kname, X, Y, bool(int(biased)), dtype=np.float64, **kw)
        assert abs(var - gvar) <= 5e-2 * abs(var6) + 1e-5, k
        n += 1
    assert n >= 10


def test_kxy_only_matches_reference():
    for k in Z.files:
        if not k.startswith("kxy|"):
            continue
        _, shape, kname = k.split("|")
        X, Y = Z["X_" + shape], Z["Y_" + shape]
        K = mmd_oracle.kernel_matrices(kname, X, Y, np.float32, K_XY_only=True)
        assert np.allclose(K, Z[k], rtol=2e-5, atol=2e-5), k


def test_quirk_add_dot_const_diagonal():
    # add_dot kernels report const_diagonal = sum(wts) -> unbiased keeps add_dot*|x|^2
    X, Y = Z["X_c1_64x16"], Z["Y_c1_64x16"]
    u = mmd_oracle.mmd2("mix_rq_1dot", X, Y, False, np.float64)
    b = mmd_oracle.mmd2("mix_rq_1dot", X, Y, True, np.float64)
    Kxx, Kxy, Kyy, cd = mmd_oracle.kernel_matrices("mix_rq_1dot", X, Y, np.float64)
    assert cd == 3.0
    true_unbiased = ((Kxx.sum() - np.trace(Kxx)) / (64 * 63) + (Kyy.sum() - np.trace(Kyy)) / (64 * 63)
                     - 2 * Kxy.mean())
    assert abs(u - true_unbiased) > 0.1 and abs(b - u) < 0.2


def test_gradient_closed_form_vs_finite_difference():
    rng = np.random.RandomState(7)
    X, Y = rng.randn(9, 4), rng.randn(11, 4) * 1.2
    for name in ("mix_rq", "mix_rbf", "distance", "dot", "mix_rq_1dot", "tanh_mix_rq", "rbf"):
        for biased in (False, True):
            v, gx, gy = mmd_oracle.mmd2_and_grads(name, X, Y, biased, np.float64)
            h = 1e-6
            for (i, j) in ((0, 0), (3, 2), (8, 3)):
                Xp, Xm = X.copy(), X.copy()
                Xp[i, j] += h
                Xm[i, j] -= h
                fd = (mmd_oracle.mmd2(name, Xp, Y, biased, np.float64)
                      - mmd_oracle.mmd2(name, Xm, Y, biased, np.float64)) / (2 * h)
                assert abs(fd - gx[i, j]) < 1e-6 * max(1, abs(fd)), (name, biased, i, j)


KZ = load_kid_golden()


@pytest.mark.parametrize("tag", ["small", "mid"])
def test_kid_oracle_matches_reference(tag):
    g, r = kid_codes(KZ, tag)
    ng, nr, d, m, S = [int(v) for v in KZ["meta_" + tag]]
    v, var = kid_oracle.polynomial_mmd(g[:m], r[:m])
    gv, gvar = KZ["pm_" + tag]
    assert _rel(v, gv) < 1e-4 and _rel(var, gvar) < 2e-2   # fp32 vs fp32 (summation order only)
    v64, var64 = kid_oracle.polynomial_mmd(g[:m].astype(np.float64), r[:m].astype(np.float64))
    assert _rel(v64, KZ["pm64_" + tag][0]) < 1e-10 and _rel(var64, KZ["pm64_" + tag][1]) < 1e-8
    assert _rel(kid_oracle.polynomial_mmd(g[:m], r[:m], ret_var=False), KZ["pm_novar_" + tag]) < 1e-4
    Kxx = kid_oracle.poly_kernel(g[:m], g[:m]).astype(np.float64)
    Kyy = kid_oracle.poly_kernel(r[:m], r[:m]).astype(np.float64)
    Kxy = kid_oracle.poly_kernel(g[:m], r[:m]).astype(np.float64)
    for est in ("biased", "unbiased", "u-statistic"):
        vv, vr = kid_oracle.mmd2_and_variance(Kxx, Kxy, Kyy, mmd_est=est, var_at_m=min(ng, nr))
        assert _rel(vv, KZ["est_%s_%s" % (est, tag)][0]) < 2e-4
    # subset loop: same global-RNG draw order as the reference (g first, then r)
    np.random.seed(0)
    mmds, vrs = kid_oracle.polynomial_mmd_averages(g.astype(np.float64), r.astype(np.float64), n_subsets=S,
                                                   subset_size=m, ret_var=True)
    assert np.allclose(mmds, KZ["avg_mmds64_" + tag], rtol=1e-9, atol=1e-14)
    assert np.allclose(vrs, KZ["avg_vars64_" + tag], rtol=1e-7, atol=1e-16)
    np.random.seed(0)
    mm32 = kid_oracle.polynomial_mmd_averages(g, r, n_subsets=S, subset_size=m, ret_var=False)
    assert np.allclose(mm32, KZ["avg_novar_" + tag], rtol=2e-3, atol=2e-6)


def test_kid_rejects_non_square():
    K = np.ones((4, 4))
    with pytest.raises(AssertionError):
        kid_oracle.mmd2_and_variance(K, np.ones((4, 5)), K)


def test_cpu_port_matches_reference_golden():
    """oracle/cpu_port.py (the bench's fallback CPU arm when the reference sources are not staged) against the
    reference-minted golden values and gradients of the mix_rq cases."""
    from oracle import cpu_port

    n = 0
    for case in INDEX:
        if case["kernel"] != "mix_rq":
            continue
        kw = dict(case["kwargs"])
        X, Y = Z["X_" + case["shape"]].astype(np.float32), Z["Y_" + case["shape"]].astype(np.float32)
        v, gx, gy = cpu_port.mix_rq_fwd_bwd(X, Y, biased=case["biased"], **kw)
        key = case["key"]
        assert abs(v - float(Z[key + "|v32"])) <= 1e-5 * max(1.0, abs(float(Z[key + "|v32"]))), key
        gx64, gy64 = Z[key + "|gx64"], Z[key + "|gy64"]
        # fp32 Gram-form distances (as in the reference): up to ~3e-5 of max|g| away from the fp64 gradients at d = 1
        assert np.abs(gx - gx64).max() <= 1e-4 * np.abs(gx64).max() + 1e-12, key
        assert np.abs(gy - gy64).max() <= 1e-4 * np.abs(gy64).max() + 1e-12, key
        n += 1
    assert n >= 5


def test_cpu_port_matches_reference_live():
    """Same, against the reference executed live (build container / staged oracle/_ref only)."""
    from oracle import cpu_port, ref_loader

    if not ref_loader.reference_available():
        pytest.skip("reference sources not present")
    rs = np.random.RandomState(7)
    X = (rs.randn(300, 48) / 7).astype(np.float32)
    Y = ((1.05 * rs.randn(280, 48) + 0.1) / 7).astype(np.float32)
    v, gx, gy = cpu_port.mix_rq_fwd_bwd(X, Y)
    rv, rgx, rgy = ref_loader.reference_loss_and_grads("mix_rq", X, Y, biased=False, dtype_name="float32")
    assert abs(v - float(rv)) <= 1e-6 * max(1.0, abs(float(rv)))
    assert np.abs(gx - rgx).max() <= 1e-5 * np.abs(rgx).max()
    assert np.abs(gy - rgy).max() <= 1e-5 * np.abs(rgy).max()
