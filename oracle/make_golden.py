"""Mint the golden fixtures under tests/golden/ by running the UNMODIFIED reference here.

Run in the build container only (needs /root/reference):   python oracle/make_golden.py
Outputs (small, committed):
  tests/golden/mmd_golden.npz  -- inputs + reference outputs of mmd2(kernel(X,Y)) and autograd
                                  gradients, fp32 (faithful TF-CPU semantics) and fp64 (truth),
                                  for every kernel name x {biased, unbiased} x several shapes;
                                  plus mmd2_and_ratio and K_XY_only cases.
  tests/golden/kid_golden.npz  -- inputs + reference outputs of polynomial_mmd /
                                  polynomial_mmd_averages (compute_scores.py) on synthetic codes.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402

KERNELS = [
    ("mix_rbf", {"sigmas": [1.0, 2.0, 4.0, 8.0, 16.0]}),
    ("mix_rbf", {}),
    ("rbf", {}),
    ("rbf", {"sigma": 2.5, "wt": 0.5}),
    ("mix_rq", {}),
    ("mix_rq", {"alphas": [0.2, 0.5, 1.0, 2.0, 5.0], "wts": [1.0, 0.5, 2.0, 1.0, 0.25]}),
    ("mix_rq_dot", {}),
    ("mix_rq_1dot", {}),
    ("mix_rq_10dot", {}),
    ("mix_rq_01dot", {}),
    ("mix_rq_001dot", {}),
    ("tanh_mix_rq", {}),
    ("distance", {}),
    ("tanh_distance", {}),
    ("dot", {}),
]

# (tag, m, n, d, scale): C1 64+64x16, the yml shape 64x1, ragged 100x7, m != n, a mid-size one
SHAPES = [
    ("c1_64x16", 64, 64, 16, 1.0),
    ("yml_64x1", 64, 64, 1, 1.0),
    ("ragged_100x7", 100, 100, 7, 1.0),
    ("mneq_48_80x5", 48, 80, 5, 1.0),
    ("mid_160x24", 160, 160, 24, 0.2),
]


def synth(m, n, d, scale, seed=1234):
    # X(fake) = N(0,1) seed 1234; Y(real) = 1.1*N(0,1)+0.1 seed 1235
    X = (np.random.RandomState(seed).randn(m, d) * scale).astype(np.float32)
    Y = ((1.1 * np.random.RandomState(seed + 1).randn(n, d) + 0.1) * scale).astype(np.float32)
    return X, Y


def make_mmd_golden(path):
    out = {}
    index = []
    mod32 = ref_loader.load_reference_mmd("float32")
    import torch

    for tag, m, n, d, scale in SHAPES:
        X, Y = synth(m, n, d, scale)
        out["X_%s" % tag] = X
        out["Y_%s" % tag] = Y
        for ki, (kname, kw) in enumerate(KERNELS):
            for biased in (False, True):
                key = "%s|%s|%d|%d" % (tag, kname, ki, int(biased))
                v32, gx32, gy32 = ref_loader.reference_loss_and_grads(kname, X, Y, biased, "float32", **kw)
                v64, gx64, gy64 = ref_loader.reference_loss_and_grads(kname, X, Y, biased, "float64", **kw)
                out[key + "|v32"] = np.float32(v32)
                out[key + "|v64"] = np.float64(v64)
                # fp64 truth gradients; rounded to fp32 storage off the C1 shape to keep fixtures small
                gdt = np.float64 if tag == "c1_64x16" else np.float32
                out[key + "|gx64"] = gx64.astype(gdt)
                out[key + "|gy64"] = gy64.astype(gdt)
                if tag == "c1_64x16":
                    out[key + "|gx32"] = gx32
                    out[key + "|gy32"] = gy32
                index.append({"key": key, "shape": tag, "kernel": kname, "kwargs": kw, "biased": biased})
        # mmd2_and_ratio (mmd.py:223) + K_XY_only on the square shapes
        if m == n:
            for kname, kw in (("mix_rq", {}), ("mix_rbf", {"sigmas": [1.0, 2.0, 4.0, 8.0, 16.0]}), ("dot", {}),
                              ("distance", {}), ("mix_rq_1dot", {})):
                for biased in (False, True):
                    K = getattr(mod32, "_%s_kernel" % kname)(torch.tensor(X), torch.tensor(Y), **kw)
                    v, r, var = mod32.mmd2_and_ratio(K, biased=biased)
                    key = "ratio|%s|%s|%d" % (tag, kname, int(biased))
                    out[key] = np.array([float(v), float(r), float(var)], dtype=np.float64)
        for kname, kw in (("mix_rq", {}), ("rbf", {}), ("distance", {}), ("dot", {}), ("mix_rq_1dot", {})):
            Kxy = getattr(mod32, "_%s_kernel" % kname)(torch.tensor(X), torch.tensor(Y), K_XY_only=True, **kw)
            out["kxy|%s|%s" % (tag, kname)] = Kxy.numpy()
    out["index_json"] = np.frombuffer(json.dumps(index).encode(), dtype=np.uint8)
    np.savez_compressed(path, **out)
    return len(index)


def make_kid_golden(path):
    cs = ref_loader.load_reference_compute_scores()
    out = {}
    # synthetic "Inception" codes: relu(N(mu,1)); small enough to commit (SURVEY 8d uses 50k x 2048)
    for tag, ng, nr, d, m in (("small", 400, 500, 64, 100), ("mid", 600, 600, 256, 250)):
        g = np.maximum(np.random.RandomState(1234).randn(ng, d), 0).astype(np.float32)
        r = np.maximum(np.random.RandomState(1235).randn(nr, d) + 0.02, 0).astype(np.float32)
        if tag == "mid":
            # do not commit 600x256 floats twice: regenerate from seeds in the test instead
            out["seeds_%s" % tag] = np.array([1234, 1235, ng, nr, d], dtype=np.int64)
        else:
            out["g_%s" % tag] = g
            out["r_%s" % tag] = r
        # one-shot polynomial_mmd on the head rows (fp32 codes -> fp32 kernel, like the scorer)
        v, var = cs.polynomial_mmd(g[:m], r[:m])
        out["pm_%s" % tag] = np.array([v, var], dtype=np.float64)
        v64, var64 = cs.polynomial_mmd(g[:m].astype(np.float64), r[:m].astype(np.float64))
        out["pm64_%s" % tag] = np.array([v64, var64], dtype=np.float64)
        v_novar = cs.polynomial_mmd(g[:m], r[:m], ret_var=False)
        out["pm_novar_%s" % tag] = np.float64(v_novar)
        # estimator variants straight on kernel blocks
        from sklearn.metrics.pairwise import polynomial_kernel as pk
        Kxx, Kyy, Kxy = pk(g[:m]), pk(r[:m]), pk(g[:m], r[:m])
        for est in ("biased", "unbiased", "u-statistic"):
            vv, vr = cs._mmd2_and_variance(Kxx, Kxy, Kyy, mmd_est=est, var_at_m=min(ng, nr))
            out["est_%s_%s" % (est, tag)] = np.array([vv, vr], dtype=np.float64)
        # subset averages with the reference's own global-RNG draw order
        np.random.seed(0)
        import io
        mmds, vrs = cs.polynomial_mmd_averages(g, r, n_subsets=6, subset_size=m, ret_var=True,
                                               output=io.StringIO())
        out["avg_mmds_%s" % tag] = mmds
        out["avg_vars_%s" % tag] = vrs
        np.random.seed(0)
        mm64, vr64 = cs.polynomial_mmd_averages(g.astype(np.float64), r.astype(np.float64), n_subsets=6,
                                                subset_size=m, ret_var=True, output=io.StringIO())
        out["avg_mmds64_%s" % tag] = mm64
        out["avg_vars64_%s" % tag] = vr64
        np.random.seed(0)
        out["avg_novar_%s" % tag] = cs.polynomial_mmd_averages(g, r, n_subsets=6, subset_size=m, ret_var=False,
                                                               output=io.StringIO())
        out["meta_%s" % tag] = np.array([ng, nr, d, m, 6], dtype=np.int64)
    np.savez_compressed(path, **out)


def fid_codes(tag):
    """Seeded synthetic "Inception" codes / class probabilities of the FID fixtures (regenerated by the tests)."""
    ng, nr, d = {"a": (900, 960, 48), "b": (1500, 1500, 96)}[tag]
    g = np.maximum(np.random.RandomState(4321).randn(ng, d) * np.linspace(0.5, 1.5, d), 0).astype(np.float32)
    r = np.maximum(np.random.RandomState(4322).randn(nr, d) * np.linspace(0.6, 1.4, d) + 0.05, 0).astype(np.float32)
    logits = np.random.RandomState(4323).randn(ng, 10).astype(np.float64) * 2.0
    p = np.exp(logits - logits.max(1, keepdims=True))
    return g, r, (p / p.sum(1, keepdims=True)).astype(np.float32)


def make_fid_golden(path):
    """Reference fid_score / inception_score (compute_scores.py:158-208) on seeded codes: 'openai' splits and 'bootstrap'
    splits (numpy global RNG seeded with 0 right before the call)."""
    import io
    import warnings

    cs = ref_loader.load_reference_compute_scores()
    out = {}
    for tag in ("a", "b"):
        g, r, p = fid_codes(tag)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")       # ComplexWarning when the reference stores a complex trace
            out["fid_openai_%s" % tag] = cs.fid_score(g, r, output=io.StringIO(), splits=3)
            out["fid64_openai_%s" % tag] = cs.fid_score(g.astype(np.float64), r.astype(np.float64), output=io.StringIO(), splits=3)
            np.random.seed(0)
            out["fid_bootstrap_%s" % tag] = cs.fid_score(g, r, output=io.StringIO(), splits=2, split_method="bootstrap")
        out["is_openai_%s" % tag] = cs.inception_score(p, splits=4)
        np.random.seed(0)
        out["is_bootstrap_%s" % tag] = cs.inception_score(p, splits=3, split_method="bootstrap")
    np.savez_compressed(path, **out)


def make_three_sample_golden(path):
    """Reference numpy 3-sample test (mmd.py:429-539) on seeded synthetic codes; fp32 (as the scorer feeds it) and
    fp64.  Inputs are regenerated from the seeds by the tests (tests/golden_util.py:three_sample_codes)."""
    ref = ref_loader.load_reference_mmd("float64")
    out = {}
    for tag, m, d in (("small", 96, 40), ("mid", 320, 192)):
        rng = np.random.RandomState(4321)
        X = np.maximum(rng.randn(m, d), 0)
        Y = np.maximum(rng.randn(m, d) + 0.05, 0)
        Z = np.maximum(rng.randn(m, d) + 0.15, 0)
        out["meta_%s" % tag] = np.array([4321, m, d], dtype=np.int64)
        for dt, dn in ((np.float32, "f32"), (np.float64, "f64")):
            x, y, z = X.astype(dt), Y.astype(dt), Z.astype(dt)
            zs = ref.np_diff_polynomial_mmd2_and_ratio_with_saving(x, z, None)
            diff, ratio, ys = ref.np_diff_polynomial_mmd2_and_ratio_with_saving(x, y, zs)
            for name, sums in (("ys", ys), ("zs", zs)):
                out["%s_%s_%s_vec" % (name, dn, tag)] = np.stack([sums[0], sums[2], sums[3]]).astype(np.float64)
                out["%s_%s_%s_sc" % (name, dn, tag)] = np.array([sums[1], sums[4]], dtype=np.float64)
            out["res_%s_%s" % (dn, tag)] = np.array([diff, ratio], dtype=np.float64)
            if dt is np.float64:
                # the graph variant (mmd.py:296-304 over the torch-backed tf shim): same sums, but the ratio divides
                # by mysqrt(max(var, eps)) = sqrt(max(var, eps) + eps) (mmd.py:398 with :12)
                import torch
                tdiff, tratio = ref.diff_polynomial_mmd2_and_ratio(torch.tensor(x), torch.tensor(y), torch.tensor(z))
                out["res_graph_%s_%s" % (dn, tag)] = np.array([float(tdiff), float(tratio)], dtype=np.float64)
    np.savez_compressed(path, **out)


if __name__ == "__main__":
    if not ref_loader.reference_available():
        sys.exit("reference not found at %s" % ref_loader.REFERENCE_ROOT)
    gdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gdir, exist_ok=True)
    if "--fid" in sys.argv:   # only the FID / inception-score fixture
        make_fid_golden(os.path.join(gdir, "fid_golden.npz"))
        print("wrote FID fixtures to %s" % gdir)
        sys.exit(0)
    if "--three-sample" in sys.argv:   # only the 3-sample fixture (leaves the committed mmd/kid files untouched)
        make_three_sample_golden(os.path.join(gdir, "three_sample_golden.npz"))
        print("wrote 3-sample fixtures to %s" % gdir)
        sys.exit(0)
    n = make_mmd_golden(os.path.join(gdir, "mmd_golden.npz"))
    make_kid_golden(os.path.join(gdir, "kid_golden.npz"))
    make_three_sample_golden(os.path.join(gdir, "three_sample_golden.npz"))
    make_fid_golden(os.path.join(gdir, "fid_golden.npz"))
    print("wrote %d mmd cases + kid fixtures to %s" % (n, gdir))
