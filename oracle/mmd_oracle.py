"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the MMD^2 loss path (numpy).

A restatement (not a copy) of the algorithm in the reference's ``gan/core/mmd.py``; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it.  The product path
(``scaled-mmd-gan_b200/smmd``) never does: it fails loudly when ``libsmmd.so`` is missing.

Pinning: the reference ships no tests/golden vectors.  This oracle is pinned
against outputs of the reference itself, executed unmodified in the build container through
``oracle/ref_loader.py`` and frozen in ``tests/golden/mmd_golden.npz`` (generator:
``oracle/make_golden.py``); ``tests/test_oracle_cpu.py`` checks oracle == golden.

Reference map (file:line relative to /root/reference):
  * pair_blocks / sq-norms from the Gram diagonal ........ gan/core/mmd.py:19-24,57-62,89-94,151-156
  * squared distances, clamp at 0 ........................ gan/core/mmd.py:67,75-76,101,109-110,163,174-175
  * distance kernel, eps *inside* the sqrt, D unclamped .. gan/core/mmd.py:6,12,29,34-35
  * dot kernel ........................................... gan/core/mmd.py:44-52
  * rbf / mix_rbf ........................................ gan/core/mmd.py:55-82 / 85-116
  * mix_rq (+add_dot) and the *_dot wrappers, tanh_* ..... gan/core/mmd.py:119-188, 40-41, 139-140
  * mmd2 biased / unbiased, const-diagonal trace ......... gan/core/mmd.py:194-220
  * mmd2_and_ratio / variance estimate ................... gan/core/mmd.py:223-293, gan/core/ops.py:209-225
  * gradient: the reference uses TF autodiff (model.py:446,452); here the closed form of
    SURVEY.md A.2, validated against autograd through the reference in the golden fixtures.
"""
from __future__ import annotations

import numpy as np

EPS = 1.0e-5  # mmd.py:6

# name -> (base family, preset kwargs).  Mirrors the `_<name>_kernel` zoo of mmd.py:18-188.
KERNEL_ZOO = {
    "distance": ("distance", {}),
    "tanh_distance": ("distance", {"tanh": True}),
    "dot": ("dot", {}),
    "rbf": ("rbf", {}),
    "mix_rbf": ("mix_rbf", {}),
    "mix_rq": ("mix_rq", {}),
    "mix_rq_dot": ("mix_rq", {"add_dot": 0.1}),
    "mix_rq_1dot": ("mix_rq", {"add_dot": 1.0}),
    "mix_rq_10dot": ("mix_rq", {"add_dot": 10.0}),
    "mix_rq_01dot": ("mix_rq", {"add_dot": 0.1}),
    "mix_rq_001dot": ("mix_rq", {"add_dot": 0.01}),
    "tanh_mix_rq": ("mix_rq", {"tanh": True}),
}

DEFAULT_SIGMAS = (2.0, 5.0, 10.0, 20.0, 40.0, 80.0)  # mmd.py:85
DEFAULT_ALPHAS = (0.1, 1.0, 10.0)  # mmd.py:143


class KernelSpec:
    """Resolved kernel description shared by every oracle entry point."""

    def __init__(self, name, sigma=1.0, wt=1.0, sigmas=None, alphas=None, wts=None, add_dot=None):
        if name not in KERNEL_ZOO:
            raise KeyError("unknown kernel %r" % (name,))
        family, preset = KERNEL_ZOO[name]
        self.name, self.family = name, family
        self.tanh = bool(preset.get("tanh", False))
        self.add_dot = float(preset.get("add_dot", 0.0) if add_dot is None else add_dot)
        if family == "rbf":
            self.params, self.wts = [float(sigma)], [float(wt)]
        elif family == "mix_rbf":
            self.params = [float(s) for s in (DEFAULT_SIGMAS if sigmas is None else sigmas)]
            self.wts = [1.0] * len(self.params) if wts is None else [float(w) for w in wts]
        elif family == "mix_rq":
            self.params = [float(a) for a in (DEFAULT_ALPHAS if alphas is None else alphas)]
            self.wts = [1.0] * len(self.params) if wts is None else [float(w) for w in wts]
        else:
            self.params, self.wts = [], []
        # what the reference returns as 4th tuple element (False -> trace taken from the matrix)
        if family in ("distance", "dot"):
            self.const_diagonal = False
        else:
            self.const_diagonal = float(sum(self.wts))  # quirk A.3-1: add_dot is NOT included


# ----------------------------------------------------------------------------------------------
# transforms on (G, n_row, n_col): value k and the two partials dk/dD, dk/dG
# ----------------------------------------------------------------------------------------------
def _pair_transform(spec, G, nr, nc, dt):
    """Returns (K, dK/dD_raw, dK/dG_direct, extra) for one block.

    D_raw = nr[:,None] + nc[None,:] - 2G.  `dK/dG_direct` is the part of the derivative that does not
    flow through D (add_dot / dot kernel).  For the distance kernel the sqrt(n) terms carry their own
    derivative w.r.t. the norms, returned in `extra` as (dK/dnr, dK/dnc) per element factors.
    """
    Draw = (-2 * G + nr[:, None] + nc[None, :]).astype(dt)
    fam = spec.family
    if fam == "dot":
        return G.astype(dt), np.zeros_like(Draw), np.ones_like(Draw), None
    if fam == "distance":
        eps = dt(EPS)
        root = np.sqrt(np.maximum(Draw + eps, 0))
        rr = np.sqrt(np.maximum(nr + eps, 0))
        rc = np.sqrt(np.maximum(nc + eps, 0))
        K = rr[:, None] + rc[None, :] - root
        with np.errstate(divide="ignore", invalid="ignore"):
            dD = np.where(Draw + eps > 0, -0.5 / root, 0.0).astype(dt)
        return K.astype(dt), dD, np.zeros_like(Draw), (0.5 / rr, 0.5 / rc)
    D = np.maximum(Draw, 0)
    live = (Draw > 0).astype(dt)  # derivative of the clamp (TF: maximum passes grad where x > 0)
    K = np.zeros_like(D)
    dD = np.zeros_like(D)
    if fam in ("rbf", "mix_rbf"):
        for s, w in zip(spec.params, spec.wts):
            gam = dt(1.0 / (2.0 * s * s))
            e = dt(w) * np.exp(-gam * D)
            K += e
            dD += -gam * e
    elif fam == "mix_rq":
        for a, w in zip(spec.params, spec.wts):
            a = dt(a)
            base = 1 + D / (2 * a)
            lg = np.log(base)
            e = dt(w) * np.exp(-a * lg)
            K += e
            dD += -0.5 * e / base
    dG = np.zeros_like(D)
    if spec.add_dot > 0:
        K = K + dt(spec.add_dot) * G
        dG = dG + dt(spec.add_dot)
    return K.astype(dt), (dD * live).astype(dt), dG, None


def _features(spec, X, Y, dt):
    X = np.asarray(X, dtype=dt)
    Y = np.asarray(Y, dtype=dt)
    if spec.tanh:
        return np.tanh(X), np.tanh(Y), X, Y
    return X, Y, X, Y


def kernel_matrices(name, X, Y, dtype=np.float32, K_XY_only=False, **kw):
    """(K_XX, K_XY, K_YY, const_diagonal) like ``mmd._<name>_kernel(X, Y, ...)``."""
    dt = np.dtype(dtype).type
    spec = KernelSpec(name, **kw)
    Xf, Yf, _, _ = _features(spec, X, Y, dt)
    XX, XY, YY = Xf @ Xf.T, Xf @ Yf.T, Yf @ Yf.T
    nx, ny = np.diagonal(XX).copy(), np.diagonal(YY).copy()
    K_XY = _pair_transform(spec, XY, nx, ny, dt)[0]
    if K_XY_only:
        return K_XY
    K_XX = _pair_transform(spec, XX, nx, nx, dt)[0]
    K_YY = _pair_transform(spec, YY, ny, ny, dt)[0]
    return K_XX, K_XY, K_YY, spec.const_diagonal


def mmd2_from_blocks(K_XX, K_XY, K_YY, const_diagonal=False, biased=False):
    """mmd.py:194-220."""
    dt = K_XX.dtype.type
    m, n = dt(K_XX.shape[0]), dt(K_YY.shape[0])
    if biased:
        return K_XX.sum() / (m * m) + K_YY.sum() / (n * n) - 2 * K_XY.sum() / (m * n)
    if const_diagonal is not False:
        tX, tY = m * dt(const_diagonal), n * dt(const_diagonal)
    else:
        tX, tY = np.trace(K_XX), np.trace(K_YY)
    return ((K_XX.sum() - tX) / (m * (m - 1)) + (K_YY.sum() - tY) / (n * (n - 1))
            - 2 * K_XY.sum() / (m * n))


def mmd2(name, X, Y, biased=False, dtype=np.float32, **kw):
    Kxx, Kxy, Kyy, cd = kernel_matrices(name, X, Y, dtype=dtype, **kw)
    return mmd2_from_blocks(Kxx, Kxy, Kyy, cd, biased)


def mmd2_and_grads(name, X, Y, biased=False, dtype=np.float64, **kw):
    """Value and closed-form gradients (dMMD2/dX, dMMD2/dY).

    Every block contributes  a * sum_ij K_ij  with a = 1/m^2 | 1/(m(m-1)) | -2/(mn); constant traces
    have zero gradient and matrix traces (distance/dot, unbiased) are removed by zeroing the diagonal
    weight.  Through D:  dD_ij/dx_i = 2(x_i - y_j);  through G: dG_ij/dx_i = y_j;  the distance
    kernel's sqrt(n) terms add  dn_i/dx_i = 2 x_i.
    """
    dt = np.dtype(dtype).type
    spec = KernelSpec(name, **kw)
    Xf, Yf, Xraw, Yraw = _features(spec, X, Y, dt)
    m, n = Xf.shape[0], Yf.shape[0]
    XX, XY, YY = Xf @ Xf.T, Xf @ Yf.T, Yf @ Yf.T
    nx, ny = np.diagonal(XX).copy(), np.diagonal(YY).copy()
    if biased:
        a_xx, a_yy = dt(1.0) / dt(m * m), dt(1.0) / dt(n * n)
    else:
        a_xx, a_yy = dt(1.0) / dt(m * (m - 1)), dt(1.0) / dt(n * (n - 1))
    a_xy = dt(-2.0) / dt(m * n)

    gX = np.zeros_like(Xf)
    gY = np.zeros_like(Yf)
    total = dt(0)
    blocks = (
        (XX, nx, nx, Xf, Xf, a_xx, True, "xx"),
        (YY, ny, ny, Yf, Yf, a_yy, True, "yy"),
        (XY, nx, ny, Xf, Yf, a_xy, False, "xy"),
    )
    for G, nr, nc, R, C, a, sym, tag in blocks:
        K, dD, dG, extra = _pair_transform(spec, G, nr, nc, dt)
        wgt = np.full_like(K, a)
        if sym and not biased and spec.const_diagonal is False:
            np.fill_diagonal(wgt, 0)  # (sum - trace): diagonal never enters
        total += (wgt * K).sum()
        if sym and not biased and spec.const_diagonal is not False:
            total -= a * dt(K.shape[0]) * dt(spec.const_diagonal)
        Wd = wgt * dD
        Wg = wgt * dG
        # rows
        gr = 2 * (Wd.sum(1)[:, None] * R - Wd @ C) + Wg @ C
        gc = 2 * (Wd.sum(0)[:, None] * C - Wd.T @ R) + Wg.T @ R
        if extra is not None:
            fr, fc = extra
            gr += (wgt.sum(1) * fr)[:, None] * 2 * R
            gc += (wgt.sum(0) * fc)[:, None] * 2 * C
        if tag == "xx":
            gX += gr + gc
        elif tag == "yy":
            gY += gr + gc
        else:
            gX += gr
            gY += gc
    if spec.tanh:
        gX = gX * (1 - Xf * Xf)
        gY = gY * (1 - Yf * Yf)
    return total, gX, gY


def mmd2_and_ratio(name, X, Y, biased=False, min_var_est=EPS, dtype=np.float32, **kw):
    """mmd.py:223-293 (incl. the unbiased-branch quirk that keeps the diagonal)."""
    Kxx, Kxy, Kyy, cd = kernel_matrices(name, X, Y, dtype=dtype, **kw)
    dt = Kxx.dtype.type
    m = dt(Kxx.shape[0])
    if cd is not False:
        dX = dY = dt(cd)
        sdX = sdY = m * dt(cd)
        sd2X = sd2Y = m * dt(cd) ** 2
    else:
        dX, dY = np.diagonal(Kxx), np.diagonal(Kyy)
        sdX, sdY = dX.sum(), dY.sum()
        sd2X, sd2Y = (dX ** 2).sum(), (dY ** 2).sum()
    rX = Kxx.sum(1) - dX
    rY = Kyy.sum(1) - dY
    c0 = Kxy.sum(0)
    c1 = Kxy.sum(1)
    sX, sY, sXY = rX.sum(), rY.sum(), c0.sum()
    qX = (Kxx ** 2).sum() - sd2X
    qY = (Kyy ** 2).sum() - sd2Y
    qXY = (Kxy ** 2).sum()
    if biased:
        val = (sX + sdX) / (m * m) + (sY + sdY) / (m * m) - 2 * sXY / (m * m)
    else:
        val = (sX + sdX) / (m * (m - 1)) + (sY + sdY) / (m * (m - 1)) - 2 * sXY / (m * m)
    var = (
        2 / (m ** 2 * (m - 1) ** 2) * (2 * (rX ** 2).sum() - qX + 2 * (rY ** 2).sum() - qY)
        - (4 * m - 6) / (m ** 3 * (m - 1) ** 3) * (sX ** 2 + sY ** 2)
        + 4 * (m - 2) / (m ** 3 * (m - 1) ** 2) * ((c1 ** 2).sum() + (c0 ** 2).sum())
        - 4 * (m - 3) / (m ** 3 * (m - 1) ** 2) * qXY
        - (8 * m - 12) / (m ** 5 * (m - 1)) * sXY ** 2
        + 8 / (m ** 3 * (m - 1)) * (1 / m * (sX + sY) * sXY - rX.dot(c1) - rY.dot(c0))
    )
    ratio = val / np.sqrt(np.maximum(var, dt(min_var_est)))
    return val, ratio, var
