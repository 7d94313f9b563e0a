"""GPU: smmd.compute_scores.fid_score / inception_score (device-resident restatement of gan/compute_scores.py:158-208:
fp64 covariance GEMMs + symmetric eigendecompositions instead of scipy's Schur sqrtm) against the fixtures minted from the
reference and against the numpy/scipy oracle.  Tolerance 1e-6 relative on FID (two different fp64 algorithms for
tr sqrtm(cov_g cov_r); measured 1e-14 .. 2e-7, the latter where the reference ran in fp32), 1e-6 on the inception score
(the reference evaluates it in the fp32 of its inputs, this module in fp64)."""
import io
import os
import warnings

import numpy as np
import pytest
import torch

from golden_util import GOLDEN_DIR
from oracle import fid_oracle, make_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["a", "b"])
def test_fid_and_inception_score_vs_reference_fixtures(tag):
    from smmd import compute_scores

    z = np.load(os.path.join(GOLDEN_DIR, "fid_golden.npz"))
    g, r, p = make_golden.fid_codes(tag)
    assert np.allclose(compute_scores.fid_score(g, r, output=io.StringIO(), splits=3), z["fid64_openai_%s" % tag], rtol=1e-6)
    np.random.seed(0)
    assert np.allclose(compute_scores.fid_score(g, r, output=io.StringIO(), splits=2, split_method="bootstrap"),
                       z["fid_bootstrap_%s" % tag], rtol=1e-6)
    assert np.allclose(compute_scores.inception_score(p, splits=4), z["is_openai_%s" % tag], rtol=1e-6)
    np.random.seed(0)
    assert np.allclose(compute_scores.inception_score(p, splits=3, split_method="bootstrap"), z["is_bootstrap_%s" % tag], rtol=1e-6)
    # CUDA tensors in: same numbers
    gt, rt = torch.tensor(g, device="cuda:0"), torch.tensor(r, device="cuda:0")
    assert np.allclose(compute_scores.fid_score(gt, rt, splits=3), z["fid64_openai_%s" % tag], rtol=1e-6)


def test_fid_rank_deficient_and_wide_codes_vs_oracle():
    """fewer rows than features per split (singular covariances: where scipy's sqrtm struggles and the reference falls back
    to eps on the diagonals) and a scorer-like width"""
    from smmd import compute_scores

    rng = np.random.RandomState(11)
    g = np.maximum(rng.randn(1200, 512), 0).astype(np.float32)
    r = np.maximum(rng.randn(1200, 512) + 0.03, 0).astype(np.float32)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = fid_oracle.fid_score(g.astype(np.float64), r.astype(np.float64), splits=2)      # 600 rows x 512 features
        ref_def = fid_oracle.fid_score(g[:600].astype(np.float64), r[:600].astype(np.float64), splits=2)   # 300 < 512
    assert np.allclose(compute_scores.fid_score(g, r, splits=2), ref, rtol=1e-6)
    assert np.allclose(compute_scores.fid_score(g[:600], r[:600], splits=2), ref_def, rtol=2e-4)   # sqrtm of a singular product
