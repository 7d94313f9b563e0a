"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the FID / inception-score helpers (numpy + scipy).

Restatement of ``gan/compute_scores.py:158-208`` of the reference (get_splits, inception_score, fid_score): per split,
mean and ``np.cov`` of the code rows, the principal matrix square root of ``cov_g @ cov_r`` by
``scipy.linalg.sqrtm`` (scipy==1.0.1 in the reference's requirements.txt:9; Schur method, not vendored), and
``|mu_g - mu_r|^2 + tr(cov_g) + tr(cov_r) - 2 tr(covmean)``.  The reference stores the (possibly complex) trace into
a float array, i.e. keeps its real part; so does this restatement.

Pinned against the reference executed here (``oracle/ref_loader.py``; its ``sqrtm(..., disp=False)`` call is served by a
signature shim because the installed scipy dropped that argument): ``tests/golden/fid_golden.npz`` written by
``oracle/make_golden.py --fid``.  Only tests / smoke() / bench.py's CPU-baseline leg may import this module.
"""
from __future__ import annotations

import numpy as np
from scipy import linalg


def get_splits(n, splits=10, split_method="openai"):
    """compute_scores.py:158-165.  'bootstrap' draws from numpy's global RNG (one np.random.choice(n, n) per split)."""
    if split_method == "openai":
        return [slice(i * n // splits, (i + 1) * n // splits) for i in range(splits)]
    if split_method == "bootstrap":
        return [np.random.choice(n, n) for _ in range(splits)]
    raise ValueError("bad split_method {}".format(split_method))


def inception_score(preds, **split_args):
    """compute_scores.py:168-176: exp(mean_i KL(p(y|x_i) || p(y))) per split."""
    preds = np.asarray(preds)
    out = []
    for inds in get_splits(preds.shape[0], **split_args):
        part = preds[inds]
        kl = part * (np.log(part) - np.log(np.mean(part, 0, keepdims=True)))
        out.append(np.exp(np.mean(np.sum(kl, 1))))
    return np.array(out)


def fid_score(codes_g, codes_r, eps=1e-6, **split_args):
    """compute_scores.py:179-208 (the g splits are drawn before the r splits)."""
    codes_g, codes_r = np.asarray(codes_g), np.asarray(codes_r)
    splits_g = get_splits(codes_g.shape[0], **split_args)
    splits_r = get_splits(codes_r.shape[0], **split_args)
    assert len(splits_g) == len(splits_r)
    d = codes_g.shape[1]
    assert codes_r.shape[1] == d
    scores = np.zeros(len(splits_g))
    for i, (w_g, w_r) in enumerate(zip(splits_g, splits_r)):
        part_g, part_r = codes_g[w_g], codes_r[w_r]
        mn_g, mn_r = part_g.mean(axis=0), part_r.mean(axis=0)
        cov_g, cov_r = np.cov(part_g, rowvar=False), np.cov(part_r, rowvar=False)
        covmean = linalg.sqrtm(cov_g.dot(cov_r))
        if not np.isfinite(covmean).all():
            cov_g[range(d), range(d)] += eps
            cov_r[range(d), range(d)] += eps
            covmean = linalg.sqrtm(cov_g.dot(cov_r))
        scores[i] = np.real(np.sum((mn_g - mn_r) ** 2) + (np.trace(cov_g) + np.trace(cov_r) - 2 * np.trace(covmean)))
    return scores
