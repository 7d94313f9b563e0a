"""CPU: the FID / inception-score oracle (oracle/fid_oracle.py, restating gan/compute_scores.py:158-208) against the
fixtures minted from the reference itself (oracle/make_golden.py --fid; the reference's sqrtm(..., disp=False) call runs
through a signature shim because the installed scipy dropped that argument)."""
import warnings

import numpy as np

from oracle import fid_oracle, make_golden
from golden_util import GOLDEN_DIR


def _golden():
    import os
    return np.load(os.path.join(GOLDEN_DIR, "fid_golden.npz"))


def test_fid_oracle_matches_reference_fixtures():
    z = _golden()
    for tag in ("a", "b"):
        g, r, _ = make_golden.fid_codes(tag)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            assert np.allclose(fid_oracle.fid_score(g, r, splits=3), z["fid_openai_%s" % tag], rtol=1e-10, atol=0)
            assert np.allclose(fid_oracle.fid_score(g.astype(np.float64), r.astype(np.float64), splits=3),
                               z["fid64_openai_%s" % tag], rtol=1e-10, atol=0)
            np.random.seed(0)
            assert np.allclose(fid_oracle.fid_score(g, r, splits=2, split_method="bootstrap"),
                               z["fid_bootstrap_%s" % tag], rtol=1e-10, atol=0)


def test_inception_score_oracle_matches_reference_fixtures():
    z = _golden()
    for tag in ("a", "b"):
        _, _, p = make_golden.fid_codes(tag)
        assert np.allclose(fid_oracle.inception_score(p, splits=4), z["is_openai_%s" % tag], rtol=1e-12)
        np.random.seed(0)
        assert np.allclose(fid_oracle.inception_score(p, splits=3, split_method="bootstrap"), z["is_bootstrap_%s" % tag],
                           rtol=1e-12)


def test_trace_sqrt_identity_used_by_the_device_path():
    """tr sqrtm(A B) = sum sqrt(eig(A^(1/2) B A^(1/2))) for symmetric PSD A, B -- the identity smmd.compute_scores.fid_score
    relies on, checked here against scipy's sqrtm in numpy (the device code itself needs a GPU: tests/test_gpu_fid.py)."""
    from scipy import linalg

    rng = np.random.RandomState(3)
    for d, n in ((24, 200), (40, 30)):          # second case: rank-deficient covariances
        A = np.cov(rng.randn(n, d), rowvar=False)
        B = np.cov(rng.randn(n, d) * 1.3 + 0.1, rowvar=False)
        w, v = np.linalg.eigh(A)
        root = (v * np.sqrt(np.clip(w, 0, None))) @ v.T
        lam = np.linalg.eigvalsh(root @ B @ root)
        mine = np.sqrt(np.clip(lam, 0, None)).sum()
        ref = np.real(np.trace(linalg.sqrtm(A @ B)))
        assert abs(mine - ref) <= 1e-6 * abs(ref)
