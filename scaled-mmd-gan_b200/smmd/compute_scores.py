"""Drop-in for the KID part of the reference's ``gan/compute_scores.py`` (lines 211-335): same
function names, kwargs and RNG draw order, backed by libsmmd.so.  (``get_splits`` / ``inception_score`` / ``fid_score``,
compute_scores.py:158-208, are provided at the end of the module as device-resident library math: fp64 covariance GEMMs
and symmetric eigendecompositions on the GPU -- plain library calls, no kernel of this repo.)

``polynomial_mmd_averages`` accepts numpy arrays (like the reference: host codes in, numpy float64
arrays out -- the codes are copied to the GPU once per call) or torch CUDA tensors (device in, torch
out, no host round trip).  Subset indices come from numpy's *global* RNG in the reference's call order
(g first, then r, per subset; compute_scores.py:221-222) and are passed to the library explicitly.
All subsets run as ONE batched pass: row gather, the three Gram blocks per subset on tensor cores,
the cubic transform and every reduction of ``_mmd2_and_variance`` fused -- no m x m matrix is written.
"""
from __future__ import annotations

import ctypes as C
import sys

import numpy as np
import torch

from . import _lib
from .mmd import _as_ptr, _stream_ptr, _workspace

_default_precision = "auto"


def set_default_precision(name):
    global _default_precision
    if name not in _lib.PRECISIONS:
        raise ValueError("precision must be one of %s" % sorted(_lib.PRECISIONS))
    _default_precision = name


def _to_device_codes(a, device):
    """numpy / CPU tensor -> CUDA tensor (pinned staging); CUDA tensor passes through."""
    if isinstance(a, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(a))
    else:
        t = a
    if t.dtype not in (torch.float32, torch.bfloat16):
        t = t.to(torch.float32)  # the reference's codes are float32 (compute_scores.py:131)
    if not t.is_cuda:
        t = t.pin_memory().to(device, non_blocking=True) if torch.cuda.is_available() else t.to(device)
    if t.stride(-1) != 1:
        t = t.contiguous()
    return t


def _draw_subsets_numpy(len_g, len_r, n_subsets, subset_size):
    choice = np.random.choice
    ig = np.empty((n_subsets, subset_size), dtype=np.int32)
    ir = np.empty((n_subsets, subset_size), dtype=np.int32)
    for i in range(n_subsets):
        ig[i] = choice(len_g, subset_size, replace=False)
        ir[i] = choice(len_r, subset_size, replace=False)
    return ig, ir


def draw_subsets(len_g, len_r, n_subsets, subset_size):
    """np.random.choice(..., replace=False) in the reference's order (compute_scores.py:219-222), from numpy's GLOBAL
    RNG and leaving it in the state the reference's loop would: the draw itself runs in the library
    (smmd_draw_subsets_mt19937: numpy's legacy permutation, bit for bit, ~3x faster than the numpy loop, which is most of
    a host-codes KID call).  Anything unusual (a sample larger than the population, a non-MT19937 state) goes through
    numpy itself, so errors and results are numpy's."""
    state = np.random.get_state()
    if (state[0] != "MT19937" or subset_size > min(len_g, len_r) or n_subsets < 1 or subset_size < 1
            or max(len_g, len_r) >= 2 ** 31):
        return _draw_subsets_numpy(len_g, len_r, n_subsets, subset_size)
    key = np.ascontiguousarray(state[1], dtype=np.uint32).copy()
    pos = C.c_int32(int(state[2]))
    ig = np.empty((n_subsets, subset_size), dtype=np.int32)
    ir = np.empty((n_subsets, subset_size), dtype=np.int32)
    st = _lib.load().smmd_draw_subsets_mt19937(key.ctypes.data_as(C.c_void_p), C.byref(pos), int(len_g), int(len_r),
                                               int(n_subsets), int(subset_size), ig.ctypes.data_as(C.c_void_p),
                                               ir.ctypes.data_as(C.c_void_p))
    _lib.check(st, "smmd_draw_subsets_mt19937")
    np.random.set_state((state[0], key, int(pos.value), state[3], state[4]))
    return ig, ir


def kid_subsets(codes_g, codes_r, idx_g, idx_r, degree=3, gamma=None, coef0=1, var_at_m=None, ret_var=True,
                mmd_est="unbiased", precision=None, first_subset=0, n_local=0):
    """Thin wrapper of smmd_kid_subsets.  codes: CUDA [n, d]; idx: CUDA int32 [S, m].
    Returns (mmd2[S] f64, var[S] f64 or None) device tensors (only the local shard is written)."""
    lib = _lib.load()
    if not (codes_g.is_cuda and codes_r.is_cuda and idx_g.is_cuda and idx_r.is_cuda):
        raise RuntimeError("smmd: KID inputs must be CUDA tensors; there is no CPU fallback")
    if codes_g.dim() != 2 or codes_r.dim() != 2 or codes_g.shape[1] != codes_r.shape[1]:
        raise ValueError("codes must be 2-D with equal feature dim")
    if idx_g.shape != idx_r.shape or idx_g.dim() != 2:
        raise ValueError("idx_g / idx_r must both be [n_subsets, subset_size]")
    S, m = idx_g.shape
    p = _lib.KidProblem()
    p.n_g, p.n_r, p.d = codes_g.shape[0], codes_r.shape[0], codes_g.shape[1]
    p.ldg, p.ldr = codes_g.stride(0), codes_r.stride(0)
    p.dtype = _lib.F32 if codes_g.dtype == torch.float32 else _lib.BF16
    p.n_subsets, p.subset_size, p.degree = S, m, int(degree)
    p.gamma = -1.0 if gamma is None else float(gamma)
    p.coef0 = float(coef0)
    p.var_at_m = int(var_at_m) if var_at_m is not None else 0
    p.mmd_est = _lib.ESTIMATORS[mmd_est]
    p.ret_var = 1 if ret_var else 0
    p.precision = _lib.PRECISIONS[precision or _default_precision]
    p.first_subset, p.n_local = int(first_subset), int(n_local)
    dev = codes_g.device
    idx_g = idx_g.to(torch.int32).contiguous()
    idx_r = idx_r.to(torch.int32).contiguous()
    with torch.cuda.device(dev):
        nbytes = lib.smmd_kid_workspace_bytes(C.byref(p))
        if nbytes == 0:
            raise _lib.SmmdError(-2, "smmd_kid_workspace_bytes", "problem rejected (shape/params)")
        ws = _workspace(nbytes, dev)
        mm = torch.zeros(S, dtype=torch.float64, device=dev)
        vv = torch.zeros(S, dtype=torch.float64, device=dev) if ret_var else None
        st = lib.smmd_kid_subsets(C.byref(p), _as_ptr(codes_g), _as_ptr(codes_r), _as_ptr(idx_g), _as_ptr(idx_r),
                                  _as_ptr(mm), _as_ptr(vv), _as_ptr(ws), nbytes, _stream_ptr(dev))
        _lib.check(st, "smmd_kid_subsets")
    return mm, vv


def polynomial_mmd_averages(codes_g, codes_r, n_subsets=50, subset_size=1000, ret_var=True, output=sys.stdout,
                            precision=None, **kernel_args):
    """compute_scores.py:211-229.  `output` is accepted for signature compatibility (no progress bar)."""
    host_in = isinstance(codes_g, np.ndarray) or (isinstance(codes_g, torch.Tensor) and not codes_g.is_cuda)
    dev = codes_g.device if (isinstance(codes_g, torch.Tensor) and codes_g.is_cuda) else torch.device("cuda")
    m = min(codes_g.shape[0], codes_r.shape[0])
    g = _to_device_codes(codes_g, dev)          # (pinned host codes: the copy is asynchronous and runs under the draw)
    r = _to_device_codes(codes_r, dev)
    ig, ir = draw_subsets(len(codes_g), len(codes_r), n_subsets, subset_size)
    igd = torch.from_numpy(ig).to(dev, non_blocking=True)
    ird = torch.from_numpy(ir).to(dev, non_blocking=True)
    mm, vv = kid_subsets(g, r, igd, ird, var_at_m=m, ret_var=ret_var, precision=precision, **kernel_args)
    if host_in:
        mm = mm.cpu().numpy()
        vv = vv.cpu().numpy() if ret_var else None
    return (mm, vv) if ret_var else mm


def polynomial_mmd(codes_g, codes_r, degree=3, gamma=None, coef0=1, var_at_m=None, ret_var=True,
                   precision=None, mmd_est="unbiased"):
    """compute_scores.py:232-244: one estimate on all rows of codes_g (X) vs codes_r (Y)."""
    if codes_g.shape[0] != codes_r.shape[0]:
        raise AssertionError("polynomial_mmd needs equally many rows in codes_g and codes_r "
                             "(square kernel blocks are asserted at compute_scores.py:259-261)")
    host_in = isinstance(codes_g, np.ndarray) or (isinstance(codes_g, torch.Tensor) and not codes_g.is_cuda)
    dev = codes_g.device if (isinstance(codes_g, torch.Tensor) and codes_g.is_cuda) else torch.device("cuda")
    g = _to_device_codes(codes_g, dev)
    r = _to_device_codes(codes_r, dev)
    m = g.shape[0]
    ident = torch.arange(m, dtype=torch.int32, device=dev).unsqueeze(0)
    mm, vv = kid_subsets(g, r, ident, ident, degree=degree, gamma=gamma, coef0=coef0, var_at_m=var_at_m,
                         ret_var=ret_var, precision=precision, mmd_est=mmd_est)
    if host_in:
        return (float(mm[0].item()), float(vv[0].item())) if ret_var else float(mm[0].item())
    return (mm[0], vv[0]) if ret_var else mm[0]


def _sqn(arr):
    """compute_scores.py:247-249."""
    flat = arr.reshape(-1)
    return flat.dot(flat)


def _row_stats_from_blocks(K_XX, K_XY, K_YY, diag_X=None, diag_Y=None):
    """Dense m x m blocks -> the [2m][6] fp64 row statistics the library's estimators work on (include/smmd.h:
    smmd_kid_from_row_stats): own-set sums without the diagonal, cross sums, their squares, the diagonal, k(x_i, y_i)."""
    dx = torch.diagonal(K_XX) if diag_X is None else diag_X
    dy = torch.diagonal(K_YY) if diag_Y is None else diag_Y
    pair = torch.diagonal(K_XY)
    sx = torch.stack([K_XX.sum(1) - dx, K_XY.sum(1), (K_XX * K_XX).sum(1) - dx * dx, (K_XY * K_XY).sum(1), dx, pair], dim=1)
    sy = torch.stack([K_YY.sum(1) - dy, K_XY.sum(0), (K_YY * K_YY).sum(1) - dy * dy, (K_XY * K_XY).sum(0), dy, pair], dim=1)
    return torch.cat([sx, sy]).contiguous()


def _mmd2_and_variance(K_XX, K_XY, K_YY, unit_diagonal=False, mmd_est='unbiased', block_size=1024,
                       var_at_m=None, ret_var=True):
    """compute_scores.py:252-335 on caller-materialised dense m x m blocks (compatibility entry point; the fused path
    is polynomial_mmd / polynomial_mmd_averages).  The blocks are reduced to per-row statistics with plain fp64
    reductions on the GPU and the estimator / variance arithmetic is the library's own (smmd_kid_from_row_stats, the
    finalizer of the fused KID path)."""
    host_in = isinstance(K_XX, np.ndarray)
    dev = torch.device("cuda")
    as_t = lambda a: (torch.from_numpy(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else a).to(dev).double()
    K_XX, K_XY, K_YY = as_t(K_XX), as_t(K_XY), as_t(K_YY)
    m = K_XX.shape[0]
    assert K_XX.shape == (m, m)
    assert K_XY.shape == (m, m)
    assert K_YY.shape == (m, m)
    if mmd_est not in _lib.ESTIMATORS:
        raise AssertionError("mmd_est must be one of %s" % sorted(_lib.ESTIMATORS))
    ones = torch.ones(m, dtype=torch.float64, device=dev) if unit_diagonal else None   # :266-269: the diagonal is taken as 1
    stats = _row_stats_from_blocks(K_XX, K_XY, K_YY, ones, ones)
    out = torch.zeros(2, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        st = _lib.load().smmd_kid_from_row_stats(_as_ptr(stats), m, _lib.ESTIMATORS[mmd_est], 1 if ret_var else 0,
                                                 int(var_at_m) if var_at_m is not None else m, _as_ptr(out[0:1]),
                                                 _as_ptr(out[1:2]), _stream_ptr(dev))
    _lib.check(st, "smmd_kid_from_row_stats")
    fin = (lambda t: float(t.item())) if host_in else (lambda t: t)
    return (fin(out[0]), fin(out[1])) if ret_var else fin(out[0])


# ------------------------------------------------------------------------------------------------------------------
# compute_scores.py:158-208 -- splits, inception score, FID.  Not part of the MMD hot path: plain library math kept on
# the device (the reference spends 34-56 s per scoring pass in scipy's sqrtm of three 2048 x 2048 products).
# ------------------------------------------------------------------------------------------------------------------
def get_splits(n, splits=10, split_method='openai'):
    """compute_scores.py:158-165; 'bootstrap' draws from numpy's global RNG exactly like the reference."""
    if split_method == 'openai':
        return [slice(i * n // splits, (i + 1) * n // splits) for i in range(splits)]
    elif split_method == 'bootstrap':
        return [np.random.choice(n, n) for _ in range(splits)]
    else:
        raise ValueError("bad split_method {}".format(split_method))


def _score_device(device):
    if not torch.cuda.is_available():
        raise RuntimeError("smmd.compute_scores needs a CUDA device (there is no CPU path in this package)")
    return torch.device(device if device is not None else "cuda")


def _take(t, inds):
    if isinstance(inds, slice):
        return t[inds]
    return t.index_select(0, torch.as_tensor(np.asarray(inds), dtype=torch.int64, device=t.device))


def inception_score(preds, device=None, **split_args):
    """compute_scores.py:168-176: exp(mean_i KL(p(y|x_i) || p(y))) per split; numpy float64 array out."""
    dev = preds.device if (torch.is_tensor(preds) and preds.is_cuda) else _score_device(device)
    p = torch.as_tensor(preds).to(dev, torch.float64)
    scores = []
    for inds in get_splits(p.shape[0], **split_args):
        part = _take(p, inds)
        kl = part * (torch.log(part) - torch.log(part.mean(0, keepdim=True)))
        scores.append(torch.exp(kl.sum(1).mean()))
    return torch.stack(scores).cpu().numpy()


def _mean_cov(part):
    """mean and np.cov(part, rowvar=False) (unbiased, n - 1) in fp64 on the device: one centred Gram X_c^T X_c."""
    mn = part.mean(0)
    xc = part - mn
    return mn, (xc.t() @ xc) / (part.shape[0] - 1)


def _trace_sqrt_product(cov_g, cov_r):
    """tr sqrtm(cov_g cov_r) for symmetric PSD factors = sum_i sqrt(lambda_i(cov_g^(1/2) cov_r cov_g^(1/2))): two symmetric
    eigendecompositions instead of the reference's complex Schur form (scipy.linalg.sqrtm, compute_scores.py:199); equal to
    the real part of the trace the reference keeps."""
    w, v = torch.linalg.eigh(cov_g)
    root = (v * w.clamp_min(0).sqrt()) @ v.t()
    m = root @ cov_r @ root
    lam = torch.linalg.eigvalsh(0.5 * (m + m.t()))
    return lam.clamp_min(0).sqrt().sum()


def fid_score(codes_g, codes_r, eps=1e-6, output=sys.stdout, device=None, **split_args):
    """compute_scores.py:179-208: per split |mu_g - mu_r|^2 + tr(cov_g) + tr(cov_r) - 2 tr sqrtm(cov_g cov_r); the g splits
    are drawn before the r splits.  numpy float64 array out (one score per split)."""
    dev = codes_g.device if (torch.is_tensor(codes_g) and codes_g.is_cuda) else _score_device(device)
    g = torch.as_tensor(codes_g).to(dev)
    r = torch.as_tensor(codes_r).to(dev)
    splits_g = get_splits(g.shape[0], **split_args)
    splits_r = get_splits(r.shape[0], **split_args)
    assert len(splits_g) == len(splits_r)
    d = g.shape[1]
    assert r.shape[1] == d
    scores = []
    for w_g, w_r in zip(splits_g, splits_r):
        mn_g, cov_g = _mean_cov(_take(g, w_g).to(torch.float64))
        mn_r, cov_r = _mean_cov(_take(r, w_r).to(torch.float64))
        tr = _trace_sqrt_product(cov_g, cov_r)
        if not bool(torch.isfinite(tr)):
            cov_g = cov_g + eps * torch.eye(d, dtype=cov_g.dtype, device=dev)
            cov_r = cov_r + eps * torch.eye(d, dtype=cov_r.dtype, device=dev)
            tr = _trace_sqrt_product(cov_g, cov_r)
        scores.append(((mn_g - mn_r) ** 2).sum() + cov_g.trace() + cov_r.trace() - 2 * tr)
    return torch.stack(scores).cpu().numpy()
