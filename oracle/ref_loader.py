"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference sources from /root/reference.

This module exists only in the build container (the GPU box has no /root/reference).  It is used
by ``oracle/make_golden.py`` to mint the fixtures under ``tests/golden/`` and by the CPU tests that
pin ``oracle/mmd_oracle.py`` / ``oracle/kid_oracle.py`` against the reference itself.

How the reference is executed without TensorFlow:

* ``gan/compute_scores.py`` only touches TF inside Inception/LeNet/featurize, so it is imported
  as-is with an empty stub ``tensorflow`` module.  ``polynomial_mmd``, ``_mmd2_and_variance`` and
  ``polynomial_mmd_averages`` (compute_scores.py:211-335) then run verbatim on numpy/sklearn.
* ``gan/core/mmd.py`` does ``import tensorflow as tf`` (mmd.py:8) and ``from .ops import dot,
  sq_sum`` (mmd.py:10).  We register a fake ``tensorflow`` module backed by torch-CPU that exposes
  exactly the ~20 functions mmd.py uses, a fake package ``core`` whose ``core.ops`` holds
  ``sq_sum``/``dot`` with the semantics of gan/core/ops.py:209-225, and exec the reference file
  under the module name ``core.mmd``.  Gradients come from torch autograd through the executed
  reference code.  dtype float32 is faithful to the TF CPU path; float64 is the truth used for
  tolerances.

Nothing under the product package imports this file.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")   # oracle/make_ref.py (travels to the GPU box)


def _pick_root():
    env = os.environ.get("SMMD_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isfile(os.path.join("/root/reference", "gan", "core", "mmd.py")):
        return "/root/reference"
    return _STAGED


REFERENCE_ROOT = _pick_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "gan", "core", "mmd.py"))


# ----------------------------------------------------------------------------------------------
# torch-backed ``tf`` shim
# ----------------------------------------------------------------------------------------------
def _make_tf_shim(torch, dtype):
    tf = types.ModuleType("tensorflow")
    tf.float32 = dtype  # the reference only ever names tf.float32; we substitute the run dtype

    class _Dim(int):
        pass

    class _Shape(tuple):   # TensorShape stand-in: iterable of dims + the one method ops.py:221-222 calls
        def assert_has_rank(self, rank):
            assert len(self) == rank, "shape %s must have rank %d" % (tuple(self), rank)

    class _NameScope:      # tf.name_scope(name, default_name, values) is a no-op context manager here
        def __init__(self, *a, **k):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    def _t(x):
        if isinstance(x, torch.Tensor):
            return x
        return torch.as_tensor(x, dtype=dtype)

    if not hasattr(torch.Tensor, "get_shape"):
        torch.Tensor.get_shape = lambda self: _Shape(_Dim(s) for s in self.shape)

    def matmul(a, b, transpose_a=False, transpose_b=False):
        a, b = _t(a), _t(b)
        if transpose_a:
            a = a.transpose(-1, -2)
        if transpose_b:
            b = b.transpose(-1, -2)
        return a @ b

    def reduce_sum(x, axis=None, keep_dims=False):
        if isinstance(x, (list, tuple)):
            # tf.reduce_sum over a python list of ints/floats (mmd.py:116) keeps the list's dtype
            if all(isinstance(v, int) for v in x):
                return torch.as_tensor(sum(x))
            return torch.as_tensor(float(sum(x)), dtype=dtype)
        x = _t(x)
        if axis is None:
            return x.sum()
        return x.sum(dim=axis, keepdim=keep_dims)

    def cast(x, dt):
        if isinstance(x, torch.Tensor):
            return x.to(dt)
        return torch.as_tensor(x, dtype=dt)

    tf.matmul = matmul
    tf.diag_part = lambda x: torch.diagonal(_t(x))
    tf.expand_dims = lambda x, axis: _t(x).unsqueeze(axis)
    tf.maximum = lambda a, b: torch.maximum(_t(a), _t(b))
    tf.sqrt = lambda x: torch.sqrt(_t(x))
    tf.exp = lambda x: torch.exp(_t(x))
    tf.log = lambda x: torch.log(_t(x))
    tf.tanh = lambda x: torch.tanh(_t(x))
    tf.reduce_sum = reduce_sum
    tf.trace = lambda x: torch.diagonal(_t(x)).sum()
    tf.cast = cast
    tf.squeeze = lambda x: _t(x).squeeze()
    tf.name_scope = _NameScope
    tf.convert_to_tensor = lambda x, name=None: _t(x)
    nn = types.ModuleType("tensorflow.nn")
    nn.l2_loss = lambda t: (_t(t) ** 2).sum() / 2
    tf.nn = nn
    return tf


def load_reference_mmd(dtype_name: str = "float32"):
    """Exec /root/reference/gan/core/mmd.py over the torch shim; returns the module.

    The module's kernels take/return torch CPU tensors of dtype ``dtype_name``.
    """
    import torch

    dtype = getattr(torch, dtype_name)
    tf = _make_tf_shim(torch, dtype)

    saved = {k: sys.modules.get(k) for k in ("tensorflow", "core", "core.ops", "core.mmd")}
    try:
        sys.modules["tensorflow"] = tf
        core = types.ModuleType("core")
        core.__path__ = []  # mark as package
        ops = types.ModuleType("core.ops")
        ops.tf = tf
        src_ops = os.path.join(REFERENCE_ROOT, "gan", "core", "ops.py")
        staged_ops = os.path.join(REFERENCE_ROOT, "gan", "core", "ops_sq_sum_dot.py")
        if os.path.isfile(staged_ops) or os.path.isfile(src_ops):
            # the reference's own `sq_sum` / `dot` (gan/core/ops.py:209-225), executed verbatim over the shim
            if os.path.isfile(staged_ops):
                with open(staged_ops) as f:
                    code = f.read()
            else:
                with open(src_ops) as f:
                    code = "".join(f.readlines()[208:225])
            exec(compile(code, "gan/core/ops.py:209-225", "exec"), ops.__dict__)
        else:   # restated: sq_sum = 2*l2_loss = sum of squares; dot of two 1-D vectors
            ops.sq_sum = lambda t, name=None: 2 * tf.nn.l2_loss(t)
            ops.dot = lambda x, y, name=None: (x.reshape(1, -1) @ y.reshape(-1, 1)).squeeze()
        core.ops = ops
        sys.modules["core"] = core
        sys.modules["core.ops"] = ops
        path = os.path.join(REFERENCE_ROOT, "gan", "core", "mmd.py")
        spec = importlib.util.spec_from_file_location("core.mmd", path)
        mod = importlib.util.module_from_spec(spec)
        mod.__package__ = "core"
        sys.modules["core.mmd"] = mod
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    mod._shim_dtype = dtype
    return mod


def load_reference_compute_scores():
    """Import /root/reference/gan/compute_scores.py with a stub ``tensorflow`` (numpy path only)."""
    saved = sys.modules.get("tensorflow")
    try:
        if saved is None:
            sys.modules["tensorflow"] = types.ModuleType("tensorflow")
        path = os.path.join(REFERENCE_ROOT, "gan", "compute_scores.py")
        spec = importlib.util.spec_from_file_location("_ref_compute_scores", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if saved is None:
            sys.modules.pop("tensorflow", None)
    _patch_sqrtm(mod)
    return mod


def _patch_sqrtm(mod):
    """The reference calls ``scipy.linalg.sqrtm(A, disp=False)`` (compute_scores.py:199, scipy==1.0.1 in its
    requirements.txt:9: returns ``(sqrtm, errest)``); the scipy installed here dropped the ``disp`` argument.  The
    reference source stays untouched: its module-level ``linalg`` name is re-bound to a namespace whose ``sqrtm``
    accepts the old signature and calls the installed (same Schur-method) implementation."""
    import inspect

    import scipy.linalg as sl

    if "disp" in inspect.signature(sl.sqrtm).parameters:
        return

    def sqrtm(A, disp=True, blocksize=64):
        out = sl.sqrtm(A)
        return out if disp else (out, 0.0)

    ns = types.SimpleNamespace(**{k: getattr(sl, k) for k in dir(sl) if not k.startswith("_")})
    ns.sqrtm = sqrtm
    mod.linalg = ns


def reference_loss_and_grads(kernel_name, X, Y, biased=False, dtype_name="float32", **kernel_kwargs):
    """Run ``mmd.mmd2(mmd._<name>_kernel(X, Y, ...), biased)`` through the reference and autograd.

    X = fake/generated features (rows), Y = real features -- the reference's argument order
    (model.py:315).  Returns (mmd2, dX, dY) as numpy arrays of ``dtype_name``.
    """
    import numpy as np
    import torch

    mod = load_reference_mmd(dtype_name)
    dt = getattr(torch, dtype_name)
    Xt = torch.tensor(np.asarray(X), dtype=dt, requires_grad=True)
    Yt = torch.tensor(np.asarray(Y), dtype=dt, requires_grad=True)
    kern = getattr(mod, "_%s_kernel" % kernel_name)
    K = kern(Xt, Yt, **kernel_kwargs)
    val = mod.mmd2(K, biased=biased)
    gX, gY = torch.autograd.grad(val, [Xt, Yt])
    return val.detach().numpy(), gX.numpy(), gY.numpy()
