"""Drop-in for the reference's ``gan/core/mmd.py`` (same function names, argument meaning and
defaults), backed by libsmmd.so.  torch CUDA tensors in, torch tensors out, differentiable.

    kernel = getattr(mmd, '_%s_kernel' % config.kernel)     # model.py:314 / smmd.py:11
    loss = mmd.mmd2(kernel(G, images))                      # model.py:315-318

``_<name>_kernel(X, Y, ...)`` returns a lazy :class:`KernelHandle` that unpacks like the reference's
``(K_XX, K_XY, K_YY, const_diagonal)`` tuple (dense matrices are only built if it IS unpacked) and
that ``mmd2`` / ``mmd2_and_ratio`` recognise so the whole loss -- Gram tiles, kernel transform,
diagonal removal, block sums and both feature gradients -- runs as one fused pass in which no N x N
matrix reaches HBM.  With ``K_XY_only=True`` a dense differentiable K_XY tensor is returned, as in
the reference (model.py:336).

There is deliberately no CPU / eager-PyTorch fallback: inputs must be CUDA tensors on a B200.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

_eps = 1.0e-5  # mmd.py:6

_DEFAULT_SIGMAS = [2.0, 5.0, 10.0, 20.0, 40.0, 80.0]  # mmd.py:85
_DEFAULT_ALPHAS = [0.1, 1.0, 10.0]  # mmd.py:143

_default_precision = "auto"


def set_default_precision(name):
    """'auto' | 'fp32' (exact SIMT, rel 1e-5 tier) | 'bf16' (tcgen05, rel 1e-3 tier) | 'bf16x3'."""
    global _default_precision
    if name not in _lib.PRECISIONS:
        raise ValueError("precision must be one of %s" % sorted(_lib.PRECISIONS))
    _default_precision = name


def _as_ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _check_features(X, Y):
    if not (isinstance(X, torch.Tensor) and isinstance(Y, torch.Tensor)):
        raise TypeError("X and Y must be torch tensors")
    if X.dim() != 2 or Y.dim() != 2 or X.shape[1] != Y.shape[1]:
        raise ValueError("X and Y must be 2-D [rows, features] with equal feature dim (got %s, %s)"
                         % (tuple(X.shape), tuple(Y.shape)))
    if not (X.is_cuda and Y.is_cuda):
        raise RuntimeError("smmd: features must be CUDA tensors on a B200; there is no CPU fallback")
    if X.dtype not in (torch.float32, torch.bfloat16) or Y.dtype != X.dtype:
        raise TypeError("smmd: features must both be float32 or both bfloat16")


class KernelSpec:
    """Kernel family + parameters, mirroring the reference's kernel constructors."""

    def __init__(self, kernel_id, params=(), wts=(), add_dot=0.0, const_diagonal=False, degree=0, name=""):
        self.kernel_id = kernel_id
        self.params = [float(v) for v in params]
        self.wts = [float(v) for v in wts]
        self.add_dot = float(add_dot)
        self.const_diagonal = const_diagonal
        self.degree = int(degree)
        self.name = name
        self._key = (self.kernel_id, tuple(self.params), tuple(self.wts), self.add_dot, self.degree)
        if len(self.params) > _lib.MAX_PARAMS:
            raise ValueError("at most %d sigmas/alphas are supported" % _lib.MAX_PARAMS)

    def problem(self, m, n, d, ldx, ldy, dtype, biased=False, precision=None, rank=0, world=1):
        p = _lib.Problem()
        p.m, p.n, p.d, p.ldx, p.ldy = m, n, d, ldx, ldy
        p.dtype = _lib.F32 if dtype == torch.float32 else _lib.BF16
        p.kernel_id = self.kernel_id
        p.nparams = len(self.params)
        for i, v in enumerate(self.params):
            p.params[i] = v
        for i, v in enumerate(self.wts):
            p.wts[i] = v
        p.add_dot = self.add_dot
        p.degree = self.degree
        p.biased = 1 if biased else 0
        p.precision = _lib.PRECISIONS[precision or _default_precision]
        p.rank, p.world = rank, world
        return p


def _rows(t):
    """Row-major view with unit inner stride (copy only if needed); returns (tensor, ld)."""
    if t.stride(1) != 1 or t.stride(0) < t.shape[1]:
        t = t.contiguous()
    return t, t.stride(0)


_problem_cache = {}   # (kernel, shapes, strides, dtype, flags) -> (smmd_problem, byref, workspace bytes, address, own m, own n)


def _load_ext():
    """smmd._C: the PyTorch C++ extension wrapper of smmd_mmd2_fwd_bwd (csrc/torch_ext, built in-tree by
    __graft_entry__.build()).  Outputs, stream and workspace are handled in C++: ~10 us less host time per call than
    the ctypes path, which matters for the batch-64 shapes (10 us of GPU time).  Optional: without it the ctypes path
    below runs the same library call."""
    import importlib.util
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_C.so")
    if os.environ.get("SMMD_NO_EXT") or not os.path.isfile(path):
        return None
    try:
        _lib.load()
        spec = importlib.util.spec_from_file_location("smmd._C", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    except Exception:   # built against another torch: fall back to ctypes (same native library)
        return None


_ext = _load_ext()
_ws_cache = {}        # (device index, stream) -> uint8 workspace tensor, grown on demand


def _workspace(nbytes, dev):
    """Per-(device, stream) workspace reused across calls (the library never keeps state in it between calls).  Calls
    on one stream are ordered, so reuse is safe; another stream gets its own buffer.  Reusing it (instead of a fresh
    torch.empty per call) keeps the caching allocator from splitting a multi-GB block for a small request and then
    paying a cudaMalloc of the full size again (seen as a 50 ms step at the 36 GB workspace of N = 65536)."""
    if torch.cuda.is_current_stream_capturing():
        # CUDA-graph capture: the buffer must live in the graph's own pool for as long as the graph does
        return torch.empty(nbytes, dtype=torch.uint8, device=dev)
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        _ws_cache.pop(key, None)
        ws = None   # drop the old buffer before asking for the larger one
        ws = _ws_cache[key] = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    return ws


def release_workspaces():
    """Free the cached workspaces (e.g. after a one-off large evaluation)."""
    _ws_cache.clear()
    if _ext is not None:
        _ext.release_workspaces()


def fused_mmd2_raw(spec, X, Y, biased=False, want_grad=True, precision=None, rank=0, world=1):
    """One call of smmd_mmd2_fwd_bwd.  Returns (scalars[16] f64 device tensor, dX, dY) for the owned rows."""
    _check_features(X, Y)
    lib = _lib.load()
    Xc, ldx = _rows(X.detach())
    Yc, ldy = _rows(Y.detach())
    m, d = Xc.shape
    n = Yc.shape[0]
    # latency-bound shapes are host-bound: the problem struct and its workspace size are cached per call signature
    key = (spec._key, m, n, d, ldx, ldy, Xc.dtype, bool(biased), precision or _default_precision, rank, world,
           bool(want_grad), _lib.options_epoch)
    cached = _problem_cache.get(key)
    if cached is None:
        prob = spec.problem(m, n, d, ldx, ldy, Xc.dtype, biased, precision, rank, world)
        nbytes = lib.smmd_mmd2_workspace_bytes(C.byref(prob), 1 if want_grad else 0)
        if nbytes == 0:
            raise _lib.SmmdError(-1, "smmd_mmd2_workspace_bytes", "problem rejected (shape/params)")
        if len(_problem_cache) > 256:
            _problem_cache.clear()
        cached = _problem_cache[key] = (prob, C.byref(prob), nbytes, C.addressof(prob),
                                        m * (rank + 1) // world - m * rank // world,
                                        n * (rank + 1) // world - n * rank // world)
    prob, prob_ref, nbytes, addr, om, on = cached
    if _ext is not None:
        st, scalars, dX, dY = _ext.mmd2_fwd_bwd(addr, Xc, Yc, om, on, bool(want_grad), nbytes)
        if st != 0:
            _lib.check(st, "smmd_mmd2_fwd_bwd")
        return scalars, dX, dY
    dev = Xc.device
    switch = torch.cuda.current_device() != dev.index
    if switch:
        prev = torch.cuda.current_device()
        torch.cuda.set_device(dev)
    try:
        ws = _workspace(nbytes, dev)
        scalars = torch.empty(_lib.NUM_SCALARS, dtype=torch.float64, device=dev)
        dX = dY = None
        if want_grad:
            dX = torch.empty((om, d), dtype=torch.float32, device=dev)
            dY = torch.empty((on, d), dtype=torch.float32, device=dev)
        st = lib.smmd_mmd2_fwd_bwd(prob_ref, _as_ptr(Xc), _as_ptr(Yc), _as_ptr(scalars), _as_ptr(dX),
                                   _as_ptr(dY), _as_ptr(ws), nbytes, _stream_ptr(dev))
        if st != 0:
            _lib.check(st, "smmd_mmd2_fwd_bwd")
    finally:
        if switch:
            torch.cuda.set_device(prev)
    return scalars, dX, dY


class _FusedMMD2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, Y, spec, biased, precision):
        need = X.requires_grad or Y.requires_grad
        scalars, dX, dY = fused_mmd2_raw(spec, X, Y, biased, want_grad=need, precision=precision)
        ctx.save_for_backward(dX, dY)
        ctx.in_dtypes = (X.dtype, Y.dtype)
        return scalars[_lib.S_MMD2].to(torch.float32)

    @staticmethod
    def backward(ctx, grad_out):
        dX, dY = ctx.saved_tensors
        if dX is None:
            return None, None, None, None, None
        g = grad_out.to(torch.float32)
        return (g * dX).to(ctx.in_dtypes[0]), (g * dY).to(ctx.in_dtypes[1]), None, None, None


class KernelHandle:
    """Lazy stand-in for the reference's ``(K_XX, K_XY, K_YY, const_diagonal)`` tuple."""

    def __init__(self, spec, X, Y):
        _check_features(X, Y)
        self.spec, self.X, self.Y = spec, X, Y
        self._dense = None

    def dense(self):
        if self._dense is None:
            s, X, Y = self.spec, self.X, self.Y
            self._dense = (kernel_xy(s, X, X), kernel_xy(s, X, Y), kernel_xy(s, Y, Y), s.const_diagonal)
        return self._dense

    def __iter__(self):
        return iter(self.dense())

    def __len__(self):
        return 4

    def __getitem__(self, i):
        return self.dense()[i]


# ---- K_XY_only: dense differentiable witness block ------------------------------------------------
class _KernelXY(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, Y, spec):
        _check_features(X, Y)
        lib = _lib.load()
        Xc, ldx = _rows(X.detach().float())
        Yc, ldy = _rows(Y.detach().float())
        m, d = Xc.shape
        n = Yc.shape[0]
        prob = spec.problem(m, n, d, ldx, ldy, torch.float32, precision="fp32")
        K = torch.empty((m, n), dtype=torch.float32, device=Xc.device)
        with torch.cuda.device(Xc.device):
            st = lib.smmd_kernel_xy(C.byref(prob), _as_ptr(Xc), _as_ptr(Yc), _as_ptr(K), n, _stream_ptr(Xc.device))
        _lib.check(st, "smmd_kernel_xy")
        ctx.save_for_backward(X, Y)   # the inputs themselves: the backward is differentiable w.r.t. them
        ctx.spec = spec
        return K

    @staticmethod
    def backward(ctx, dK):
        X, Y = ctx.saved_tensors
        dX, dY = _KernelXYBackward.apply(X, Y, dK, ctx.spec)
        return dX, dY, None


class _KernelXYBackward(torch.autograd.Function):
    """(X, Y, dK) -> (dX, dY) of the witness block, itself differentiable once more: the gradient penalty
    (model.py:336-341) takes the gradient of the witness w.r.t. the critic input and then differentiates its norm
    w.r.t. the critic weights, which runs through this op's backward (smmd_kernel_xy_bwd2)."""

    @staticmethod
    def forward(ctx, X, Y, dK, spec):
        lib = _lib.load()
        Xc, _ = _rows(X.detach().float())
        Yc, _ = _rows(Y.detach().float())
        m, d = Xc.shape
        n = Yc.shape[0]
        dKc = dK.detach().contiguous().float()
        prob = spec.problem(m, n, d, Xc.stride(0), Yc.stride(0), torch.float32, precision="fp32")
        dX = torch.empty((m, d), dtype=torch.float32, device=Xc.device)
        dY = torch.empty((n, d), dtype=torch.float32, device=Xc.device)
        with torch.cuda.device(Xc.device):
            st = lib.smmd_kernel_xy_bwd(C.byref(prob), _as_ptr(Xc), _as_ptr(Yc), _as_ptr(dKc), n, _as_ptr(dX),
                                        _as_ptr(dY), _stream_ptr(Xc.device))
        _lib.check(st, "smmd_kernel_xy_bwd")
        ctx.save_for_backward(Xc, Yc, dKc)
        ctx.spec = spec
        ctx.dtypes = (X.dtype, Y.dtype, dK.dtype)
        return dX.to(X.dtype), dY.to(Y.dtype)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, VX, VY):
        Xc, Yc, dKc = ctx.saved_tensors
        lib = _lib.load()
        spec = ctx.spec
        m, d = Xc.shape
        n = Yc.shape[0]
        VXc = VX.contiguous().float() if VX is not None else None
        VYc = VY.contiguous().float() if VY is not None else None
        prob = spec.problem(m, n, d, Xc.stride(0), Yc.stride(0), torch.float32, precision="fp32")
        ddK = torch.empty((m, n), dtype=torch.float32, device=Xc.device)
        gX = torch.empty((m, d), dtype=torch.float32, device=Xc.device)
        gY = torch.empty((n, d), dtype=torch.float32, device=Xc.device)
        with torch.cuda.device(Xc.device):
            st = lib.smmd_kernel_xy_bwd2(C.byref(prob), _as_ptr(Xc), _as_ptr(Yc), _as_ptr(dKc), n, _as_ptr(VXc),
                                         _as_ptr(VYc), _as_ptr(ddK), _as_ptr(gX), _as_ptr(gY), _stream_ptr(Xc.device))
        _lib.check(st, "smmd_kernel_xy_bwd2")
        return gX.to(ctx.dtypes[0]), gY.to(ctx.dtypes[1]), ddK.to(ctx.dtypes[2]), None


_TANH_BASE = {_lib.K_TANH_DISTANCE: _lib.K_DISTANCE, _lib.K_TANH_MIX_RQ: _lib.K_MIX_RQ}


def kernel_xy(spec, X, Y):
    """Dense, twice-differentiable K_XY (K_XY_only=True).  tanh_* kernels (mmd.py:40,139) apply tanh to the
    features first, exactly as the reference does, and then use the base kernel."""
    if spec.kernel_id in _TANH_BASE:
        base = KernelSpec(_TANH_BASE[spec.kernel_id], spec.params, spec.wts, spec.add_dot, spec.const_diagonal,
                          spec.degree, spec.name)
        return _KernelXY.apply(torch.tanh(X), torch.tanh(Y), base)
    return _KernelXY.apply(X, Y, spec)


def _make(spec, X, Y, K_XY_only):
    if K_XY_only:
        return kernel_xy(spec, X, Y)
    return KernelHandle(spec, X, Y)


# ---- the kernel zoo: names, defaults and argument order of gan/core/mmd.py:18-188 -------------------
def _distance_kernel(X, Y, K_XY_only=False):
    return _make(KernelSpec(_lib.K_DISTANCE, name="distance"), X, Y, K_XY_only)


def _tanh_distance_kernel(X, Y, K_XY_only=False):
    return _make(KernelSpec(_lib.K_TANH_DISTANCE, name="tanh_distance"), X, Y, K_XY_only)


def _dot_kernel(X, Y, K_XY_only=False):
    return _make(KernelSpec(_lib.K_DOT, name="dot"), X, Y, K_XY_only)


def _rbf_kernel(X, Y, sigma=1., wt=1., K_XY_only=False):
    spec = KernelSpec(_lib.K_RBF, [sigma], [wt], const_diagonal=float(wt), name="rbf")
    return _make(spec, X, Y, K_XY_only)


def _mix_rbf_kernel(X, Y, sigmas=None, wts=None, K_XY_only=False):
    sigmas = _DEFAULT_SIGMAS if sigmas is None else list(sigmas)
    wts = [1.0] * len(sigmas) if wts is None else list(wts)
    spec = KernelSpec(_lib.K_MIX_RBF, sigmas, wts, const_diagonal=float(sum(wts)), name="mix_rbf")
    return _make(spec, X, Y, K_XY_only)


def _mix_rq_kernel(X, Y, alphas=None, wts=None, K_XY_only=False, add_dot=.0, _tanh=False):
    alphas = _DEFAULT_ALPHAS if alphas is None else list(alphas)
    wts = [1.0] * len(alphas) if wts is None else list(wts)
    # const_diagonal is sum(wts) even when add_dot > 0 -- reference quirk kept (mmd.py:186-188)
    spec = KernelSpec(_lib.K_TANH_MIX_RQ if _tanh else _lib.K_MIX_RQ, alphas, wts, add_dot=add_dot,
                      const_diagonal=float(sum(wts)), name="tanh_mix_rq" if _tanh else "mix_rq")
    return _make(spec, X, Y, K_XY_only)


def _mix_rq_dot_kernel(X, Y, alphas=None, wts=None, K_XY_only=False):
    return _mix_rq_kernel(X, Y, alphas=alphas, wts=wts, K_XY_only=K_XY_only, add_dot=.1)


def _mix_rq_1dot_kernel(X, Y, alphas=None, wts=None, K_XY_only=False):
    return _mix_rq_kernel(X, Y, alphas=alphas, wts=wts, K_XY_only=K_XY_only, add_dot=1.)


def _mix_rq_10dot_kernel(X, Y, alphas=None, wts=None, K_XY_only=False):
    return _mix_rq_kernel(X, Y, alphas=alphas, wts=wts, K_XY_only=K_XY_only, add_dot=10.)


def _mix_rq_01dot_kernel(X, Y, alphas=None, wts=None, K_XY_only=False):
    return _mix_rq_kernel(X, Y, alphas=alphas, wts=wts, K_XY_only=K_XY_only, add_dot=.1)


def _mix_rq_001dot_kernel(X, Y, alphas=None, wts=None, K_XY_only=False):
    return _mix_rq_kernel(X, Y, alphas=alphas, wts=wts, K_XY_only=K_XY_only, add_dot=.01)


def _tanh_mix_rq_kernel(X, Y, K_XY_only=False):
    return _mix_rq_kernel(X, Y, K_XY_only=K_XY_only, _tanh=True)


# ---- estimators -------------------------------------------------------------------------------------
def mmd2(K, biased=False, precision=None):
    """mmd.py:194-196.  ``K`` is a KernelHandle (fused path) or an explicit 4-tuple of dense blocks."""
    if isinstance(K, KernelHandle):
        return _FusedMMD2.apply(K.X, K.Y, K.spec, bool(biased), precision)
    K_XX, K_XY, K_YY, const_diagonal = K
    return _mmd2(K_XX, K_XY, K_YY, const_diagonal, biased)


def _mmd2(K_XX, K_XY, K_YY, const_diagonal=False, biased=False):
    """mmd.py:199-220 on caller-materialised dense blocks (compatibility path: plain reductions)."""
    m = float(K_XX.shape[0])
    n = float(K_YY.shape[0])
    if biased:
        return K_XX.sum() / (m * m) + K_YY.sum() / (n * n) - 2 * K_XY.sum() / (m * n)
    if const_diagonal is not False:
        trace_X, trace_Y = m * float(const_diagonal), n * float(const_diagonal)
    else:
        trace_X, trace_Y = torch.trace(K_XX), torch.trace(K_YY)
    return ((K_XX.sum() - trace_X) / (m * (m - 1)) + (K_YY.sum() - trace_Y) / (n * (n - 1))
            - 2 * K_XY.sum() / (m * n))


def mmd2_and_ratio(K, biased=False, min_var_est=_eps):
    """mmd.py:223-232 -> (mmd2, ratio, var_est); fused statistics pass, not differentiable."""
    if not isinstance(K, KernelHandle):   # explicit dense 4-tuple, as mmd2() accepts (mmd.py:224-225)
        K_XX, K_XY, K_YY, const_diagonal = K
        return _mmd2_and_ratio(K_XX, K_XY, K_YY, const_diagonal, biased, min_var_est)
    X, Y, spec = K.X, K.Y, K.spec
    lib = _lib.load()
    Xc, ldx = _rows(X.detach())
    Yc, ldy = _rows(Y.detach())
    m, d = Xc.shape
    n = Yc.shape[0]
    if m != n:
        raise ValueError("mmd2_and_ratio assumes X and Y have the same number of rows (mmd.py:237)")
    prob = spec.problem(m, n, d, ldx, ldy, Xc.dtype, biased, "fp32")
    dev = Xc.device
    with torch.cuda.device(dev):
        # statistics run on the exact path: same workspace formula as a forward-only fp32 problem
        nbytes = lib.smmd_mmd2_workspace_bytes(C.byref(prob), 0)
        ws = _workspace(nbytes, dev)
        scalars = torch.empty(_lib.NUM_SCALARS, dtype=torch.float64, device=dev)
        st = lib.smmd_mmd2_and_ratio(C.byref(prob), _as_ptr(Xc), _as_ptr(Yc), float(min_var_est), _as_ptr(scalars),
                                     _as_ptr(ws), nbytes, _stream_ptr(dev))
        _lib.check(st, "smmd_mmd2_and_ratio")
    return (scalars[_lib.S_MMD2].float(), scalars[_lib.S_RATIO].float(), scalars[_lib.S_VAR].float())


def _ratio_scalars_from_blocks(K_XX, K_XY, K_YY, const_diagonal, biased, min_var_est):
    """Dense blocks -> row statistics (plain reductions) -> the library's ratio finalizer (smmd_ratio_from_row_stats:
    the same kernel that finishes the fused mmd2_and_ratio, reference quirks included)."""
    from .compute_scores import _row_stats_from_blocks

    m = K_XX.shape[0]
    if K_XX.shape != (m, m) or K_XY.shape != (m, m) or K_YY.shape != (m, m):
        raise ValueError("mmd2_and_ratio assumes X and Y have the same number of rows (mmd.py:237)")
    dev = K_XX.device
    stats = _row_stats_from_blocks(K_XX.double(), K_XY.double(), K_YY.double())
    scalars = torch.empty(_lib.NUM_SCALARS, dtype=torch.float64, device=dev)
    has_cd = const_diagonal is not False
    with torch.cuda.device(dev):
        st = _lib.load().smmd_ratio_from_row_stats(_as_ptr(stats), m, 1 if biased else 0, 1 if has_cd else 0,
                                                   float(const_diagonal) if has_cd else 0.0, float(min_var_est),
                                                   _as_ptr(scalars), _stream_ptr(dev))
    _lib.check(st, "smmd_ratio_from_row_stats")
    return scalars


def _mmd2_and_variance(K_XX, K_XY, K_YY, const_diagonal=False, biased=False):
    """mmd.py:236-293 on caller-materialised dense blocks (CUDA tensors): (mmd2, var_est), including the reference's
    quirk that the unbiased estimate keeps the diagonal (:273-276)."""
    sc = _ratio_scalars_from_blocks(K_XX, K_XY, K_YY, const_diagonal, biased, _eps)
    return sc[_lib.S_MMD2].to(K_XX.dtype), sc[_lib.S_VAR].to(K_XX.dtype)


def _mmd2_and_ratio(K_XX, K_XY, K_YY, const_diagonal=False, biased=False, min_var_est=_eps):
    """mmd.py:228-233 on dense blocks: (mmd2, ratio = mmd2 / sqrt(max(var_est, min_var_est)), var_est)."""
    sc = _ratio_scalars_from_blocks(K_XX, K_XY, K_YY, const_diagonal, biased, min_var_est)
    return sc[_lib.S_MMD2].to(K_XX.dtype), sc[_lib.S_RATIO].to(K_XX.dtype), sc[_lib.S_VAR].to(K_XX.dtype)


# ---------------------------------------------------------------------------------------------------
# 3-sample test of the scorer (reference gan/core/mmd.py:296-539, caller gan/utils/scorer.py:119-163)
# ---------------------------------------------------------------------------------------------------
def polynomial_related_sums(X, Y, precision=None):
    """The five sums of `_get_sums` / `_np_get_sums` (mmd.py:405-426, 515-539) for the cubic kernel
    K = (A B^T / d + 1)^3, computed on the GPU without materialising K_XY / K_YY (smmd_poly_sums).
    X, Y: CUDA [m, d] of equal shape.  Returns (Kt_YY_sums[m], Kt_YY_2_sum, K_XY_sums_0[m], K_XY_sums_1[m],
    K_XY_2_sum) as float64 device tensors."""
    lib = _lib.load()
    if not (isinstance(X, torch.Tensor) and isinstance(Y, torch.Tensor) and X.is_cuda and Y.is_cuda):
        raise RuntimeError("smmd: the 3-sample sums need CUDA tensors; there is no CPU fallback")
    if X.dim() != 2 or X.shape != Y.shape:
        raise ValueError("X and Y must be 2-D with the same shape (mmd.py:516)")
    Xc, ldx = _rows(X.detach())
    Yc, ldy = _rows(Y.detach())
    if Yc.dtype != Xc.dtype:
        Yc = Yc.to(Xc.dtype)
    m, d = Xc.shape
    p = _lib.KidProblem()
    p.n_g = p.n_r = m
    p.d = d
    p.ldg, p.ldr = ldx, ldy
    p.dtype = _lib.F32 if Xc.dtype == torch.float32 else _lib.BF16
    p.n_subsets, p.subset_size, p.degree = 1, m, 3
    p.gamma, p.coef0 = -1.0, 1.0
    p.var_at_m, p.mmd_est, p.ret_var = 0, _lib.ESTIMATORS["unbiased"], 1
    p.precision = _lib.PRECISIONS[precision or "auto"]
    p.first_subset, p.n_local = 0, 0
    dev = Xc.device
    with torch.cuda.device(dev):
        nbytes = lib.smmd_poly_sums_workspace_bytes(C.byref(p))
        if nbytes == 0:
            raise _lib.SmmdError(-2, "smmd_poly_sums_workspace_bytes", "problem rejected (shape/params)")
        ws = _workspace(nbytes, dev)
        out = torch.empty(3 * m + 2, dtype=torch.float64, device=dev)
        st = lib.smmd_poly_sums(C.byref(p), _as_ptr(Xc), _as_ptr(Yc), _as_ptr(out), _as_ptr(ws), nbytes, _stream_ptr(dev))
        _lib.check(st, "smmd_poly_sums")
    return out[:m], out[3 * m], out[m:2 * m], out[2 * m:3 * m], out[3 * m + 1]


def _diff_mmd2_and_ratio_from_sums(Y_related_sums, Z_related_sums, m, const_diagonal=False):
    """mmd.py:339-402 / 447-512: MMD^2(X,Y) - MMD^2(X,Z), and its ratio to the estimated standard deviation.
    Works on torch tensors or numpy arrays (vectors of length m and scalars); `const_diagonal` is accepted for
    signature compatibility (the reference ignores it here as well).  The two reference variants clamp differently
    and both are kept: the graph (tensor) variant divides by mysqrt(max(var, eps)) = sqrt(max(var, eps) + eps)
    (mmd.py:398 with :12), the numpy variant by sqrt(max(var, eps)) (mmd.py:510)."""
    yy_r, yy_q, xy_c, xy_r, xy_q = Y_related_sums
    zz_r, zz_q, xz_c, xz_r, xz_q = Z_related_sums
    m = float(m)
    mm1, mm = m * (m - 1), m * m
    mu_yy, mu_zz = yy_r.sum() / mm1, zz_r.sum() / mm1
    mu_xy, mu_xz = xy_c.sum() / mm, xz_c.sum() / mm
    t3, t2 = mm1 * (m - 2), mm * (m - 1)
    e_y_yy, e_z_zz = ((yy_r * yy_r).sum() - yy_q) / t3, ((zz_r * zz_r).sum() - zz_q) / t3
    e_x_xy, e_x_xz = ((xy_r * xy_r).sum() - xy_q) / t2, ((xz_r * xz_r).sum() - xz_q) / t2
    e_y_xy, e_z_xz = ((xy_c * xy_c).sum() - xy_q) / t2, ((xz_c * xz_c).sum() - xz_q) / t2
    c_y, c_z = (yy_r * xy_c).sum() / t2, (zz_r * xz_c).sum() / t2
    c_x = (xy_r * xz_r).sum() / (mm * m)
    cross = (-2 * c_y + 2 * mu_yy * mu_xy - 2 * c_x + 2 * mu_xy * mu_xz - 2 * c_z + 2 * mu_zz * mu_xz)
    mmd2_diff = mu_yy - 2 * mu_xy - mu_zz + 2 * mu_xz
    first_order = 4 * (m - 2) / mm1 * (e_y_yy - mu_yy ** 2 + e_x_xy - mu_xy ** 2 + e_y_xy - mu_xy ** 2
                                       + e_z_zz - mu_zz ** 2 + e_x_xz - mu_xz ** 2 + e_z_xz - mu_xz ** 2 + cross)
    second_order = 2 / mm1 * (yy_q / mm1 - mu_yy ** 2 + 2 * xy_q / mm - 2 * mu_xy ** 2
                              + zz_q / mm1 - mu_zz ** 2 + 2 * xz_q / mm - 2 * mu_xz ** 2 + 2 * cross)
    var_est = first_order + second_order
    if isinstance(var_est, torch.Tensor):
        ratio = mmd2_diff / torch.sqrt(torch.clamp(var_est, min=_eps) + _eps)
    else:
        ratio = mmd2_diff / max(float(var_est), _eps) ** 0.5
    return mmd2_diff, ratio


_np_diff_mmd2_and_ratio_from_sums = _diff_mmd2_and_ratio_from_sums   # mmd.py:447 (same arithmetic on numpy)


def _sums_like(sums, like):
    """Hand the sums back in the caller's world: numpy in -> numpy (float64) out."""
    if isinstance(like, torch.Tensor):
        return sums
    return tuple(s.cpu().numpy() if s.dim() else float(s.item()) for s in sums)


def _codes_to_device(a):
    if isinstance(a, torch.Tensor):
        return a if a.is_cuda else a.cuda()
    import numpy as np
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def np_diff_polynomial_mmd2_and_ratio_with_saving(X, Y, saved_sums_for_Z, precision=None):
    """mmd.py:429-444 (scorer.py:129,157): numpy (or torch) codes X = real, Y = new fake sample, and the sums
    saved from an earlier call for the older sample Z.  Returns the sums of (X, Y) when `saved_sums_for_Z` is
    None, else (mmd2_diff, ratio, Y_related_sums)."""
    sums = polynomial_related_sums(_codes_to_device(X), _codes_to_device(Y), precision=precision)
    if saved_sums_for_Z is None:
        return _sums_like(sums, X)
    m = float(Y.shape[0])
    if isinstance(X, torch.Tensor):
        saved = tuple(torch.as_tensor(s, dtype=torch.float64, device=sums[0].device) for s in saved_sums_for_Z)
        diff, ratio = _diff_mmd2_and_ratio_from_sums(sums, saved, m)
        return diff, ratio, sums
    host = _sums_like(sums, X)
    diff, ratio = _diff_mmd2_and_ratio_from_sums(host, saved_sums_for_Z, m)
    return float(diff), float(ratio), host


def diff_polynomial_mmd2_and_ratio_with_saving(X, Y, saved_sums_for_Z, precision=None):
    """mmd.py:307-319 (graph version of the above): torch tensors in, torch scalars out."""
    sums = polynomial_related_sums(X, Y, precision=precision)
    diff, ratio = _diff_mmd2_and_ratio_from_sums(sums, saved_sums_for_Z, float(Y.shape[0]))
    return diff, ratio, sums


def diff_polynomial_mmd2_and_ratio(X, Y, Z, precision=None):
    """mmd.py:296-304: MMD^2(X,Y) - MMD^2(X,Z) with the cubic kernel, and its test statistic."""
    ys = polynomial_related_sums(X, Y, precision=precision)
    zs = polynomial_related_sums(X, Z, precision=precision)
    return _diff_mmd2_and_ratio_from_sums(ys, zs, float(Y.shape[0]))


def _np_get_sums(K_XY, K_YY, const_diagonal=False):
    """mmd.py:515-539 on dense blocks (numpy or torch), for callers that already hold the matrices."""
    m = float(K_YY.shape[0])
    if const_diagonal is not False:
        diag = float(const_diagonal)
        diag2 = m * diag ** 2
    else:
        diag = K_YY.diagonal()
        diag2 = (diag * diag).sum()
    return (K_YY.sum(1) - diag, (K_YY * K_YY).sum() - diag2, K_XY.sum(0), K_XY.sum(1), (K_XY * K_XY).sum())


_get_sums = _np_get_sums   # mmd.py:405-426


def _diff_mmd2_and_ratio(K_XY, K_XZ, K_YY, K_ZZ, const_diagonal=False):
    """mmd.py:322-336 on dense blocks the caller already holds (numpy or torch): MMD^2(X,Y) - MMD^2(X,Z) and ratio."""
    m = float(K_YY.shape[0])
    return _diff_mmd2_and_ratio_from_sums(_np_get_sums(K_XY, K_YY, const_diagonal),
                                          _np_get_sums(K_XZ, K_ZZ, const_diagonal), m, const_diagonal=const_diagonal)
