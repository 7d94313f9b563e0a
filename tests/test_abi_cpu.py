"""CPU: the C-ABI library loads, exports every symbol include/smmd.h declares, validates arguments
before touching the device, and refuses to run without a B200 (no fallback)."""
import ctypes as C
import os
import re

import pytest

from smmd import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "smmd.h")).read()
    return sorted(set(re.findall(r"SMMD_API\s+[\w\s\*]+?\b(smmd_\w+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    assert sorted(_lib.EXPORTS) == declared
    for name in declared:
        assert getattr(lib, name) is not None


def test_version_and_strerror():
    lib = _lib.load()
    assert lib.smmd_version() == 100
    assert lib.smmd_strerror(0) == b"ok"
    assert b"sm_100" in lib.smmd_strerror(-4)


def _problem(**kw):
    p = _lib.Problem()
    p.m, p.n, p.d, p.ldx, p.ldy = 64, 64, 16, 16, 16
    p.dtype, p.kernel_id, p.nparams = _lib.F32, _lib.K_MIX_RBF, 2
    p.params[0], p.params[1] = 1.0, 2.0
    p.wts[0], p.wts[1] = 1.0, 1.0
    p.precision, p.rank, p.world = _lib.PREC_FP32, 0, 1
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def test_validation_happens_before_device_access():
    lib = _lib.load()
    dummy = C.c_void_p(256)
    call = lambda p: lib.smmd_mmd2_fwd_bwd(C.byref(p), dummy, dummy, dummy, None, None, dummy, 1 << 20, None)
    assert call(_problem(m=0)) == -2            # SMMD_ESHAPE
    assert call(_problem(ldx=8)) == -2
    assert call(_problem(m=1)) == -2            # unbiased needs m >= 2
    assert call(_problem(dtype=7)) == -3        # SMMD_EDTYPE
    assert call(_problem(world=2, rank=2)) == -1
    assert call(_problem(kernel_id=99)) == -1
    assert lib.smmd_mmd2_workspace_bytes(C.byref(_problem(m=0)), 1) == 0
    assert lib.smmd_mmd2_workspace_bytes(C.byref(_problem()), 1) > 0
    # half-specified gradient outputs are rejected
    assert lib.smmd_mmd2_fwd_bwd(C.byref(_problem()), dummy, dummy, dummy, dummy, None, dummy, 1 << 20, None) == -1


def test_no_fallback_without_b200():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    assert lib.smmd_device_supported() == 0
    dummy = C.c_void_p(256)
    st = lib.smmd_mmd2_fwd_bwd(C.byref(_problem()), dummy, dummy, dummy, None, None, dummy, 1 << 20, None)
    assert st == -4  # SMMD_EARCH: fails loudly, never computes on the CPU
    from smmd import mmd

    with pytest.raises(RuntimeError):
        mmd.mmd2(mmd._rbf_kernel(torch.zeros(4, 2), torch.zeros(4, 2)))


def test_kid_validation():
    lib = _lib.load()
    p = _lib.KidProblem()
    p.n_g, p.n_r, p.d, p.ldg, p.ldr = 100, 100, 32, 32, 32
    p.n_subsets, p.subset_size, p.degree, p.coef0 = 4, 50, 3, 1.0
    p.precision = _lib.PREC_FP32
    assert lib.smmd_kid_workspace_bytes(C.byref(p)) > 0
    p.subset_size = 101  # replace=False cannot draw more rows than exist
    assert lib.smmd_kid_workspace_bytes(C.byref(p)) == 0


def test_kid_shapes_outside_the_tensor_core_kernel_are_settled_before_any_launch():
    """subset_size > 8192 (more than 64 row blocks of 256 per stacked subset) is not covered by the macro-tile kernel:
    AUTO resolves to the exact path, an explicit bf16 / bf16x3 request is refused (workspace query 0, call returns
    SMMD_EUNSUPPORTED on a B200) -- nothing is queued first."""
    lib = _lib.load()
    p = _lib.KidProblem()
    p.n_g, p.n_r, p.d, p.ldg, p.ldr = 20000, 20000, 64, 64, 64
    p.n_subsets, p.subset_size, p.degree, p.coef0 = 2, 9000, 3, 1.0
    p.precision = _lib.PREC_AUTO
    auto_bytes = lib.smmd_kid_workspace_bytes(C.byref(p))
    p.precision = _lib.PREC_FP32
    assert auto_bytes > 0 and auto_bytes == lib.smmd_kid_workspace_bytes(C.byref(p))
    for prec in (_lib.PREC_BF16, _lib.PREC_BF16X3):
        p.precision = prec
        assert lib.smmd_kid_workspace_bytes(C.byref(p)) == 0
    p.subset_size, p.precision = 8192, _lib.PREC_BF16X3
    assert lib.smmd_kid_workspace_bytes(C.byref(p)) > 0


def test_set_option_validates_names_and_clamps():
    lib = _lib.load()
    assert lib.smmd_set_option(b"no_such_option", 1) == -1
    assert lib.smmd_set_option(b"wz_min_d", 100000) == 0     # clamped to 256: the fused kernel cannot hold a wider O
    assert lib.smmd_set_option(b"wz_min_d", 256) == 0
    assert lib.smmd_set_option(b"sym_min_rows", 0) == 0
    assert lib.smmd_set_option(b"debug_nullmath", 1) == -1   # ablation knobs exist only in developer builds


def test_subset_draw_is_numpys_bit_for_bit():
    """smmd_draw_subsets_mt19937 (host helper behind compute_scores.draw_subsets): the same indices as the reference's
    np.random.choice(n, m, replace=False) loop (gan/compute_scores.py:219-222) AND the same global RNG state afterwards,
    for several population / sample sizes, stream positions and cached-gaussian states."""
    import numpy as np

    from smmd import compute_scores as cs

    cases = [(5000, 6000, 7, 300, 0), (400, 500, 6, 100, 1), (7, 9, 5, 7, 3), (1000, 1000, 3, 1, 5), (65537, 70000, 2, 33, 9),
             (1, 1, 2, 1, 1), (2, 3, 4, 2, 2), (1024, 1025, 3, 1000, 4), (50000, 50000, 2, 1000, 6),
             (50000, 49999, 6, 1000, 7), (70000, 300000, 3, 50, 8)]     # the last two take the scan + threaded-shuffle path
    for (lg, lr, S, m, seed), threads in [(c, t) for c in cases for t in ("1", "3", None)]:
        if threads is None:
            os.environ.pop("SMMD_DRAW_THREADS", None)
        else:
            os.environ["SMMD_DRAW_THREADS"] = threads
        for burn in (0, 3):                      # burn = 3 leaves a cached gaussian and an odd stream position behind
            np.random.seed(seed)
            if burn:
                np.random.randn(burn)
            a = cs._draw_subsets_numpy(lg, lr, S, m)
            sa = np.random.get_state()
            np.random.seed(seed)
            if burn:
                np.random.randn(burn)
            b = cs.draw_subsets(lg, lr, S, m)
            sb = np.random.get_state()
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (lg, lr, S, m, seed)
            assert sa[0] == sb[0] and np.array_equal(sa[1], sb[1]) and sa[2:] == sb[2:], (lg, lr, S, m, seed)
    os.environ.pop("SMMD_DRAW_THREADS", None)
    # a sample larger than the population: numpy's own error, raised by numpy
    with pytest.raises(ValueError):
        cs.draw_subsets(10, 20, 2, 11)
