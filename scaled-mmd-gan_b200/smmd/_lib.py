"""ctypes binding of include/smmd.h (libsmmd.so).  No fallbacks: if the library is missing or the
device is not a B200 the product path raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SMMD_LIB", os.path.join(os.path.dirname(_HERE), "lib", "libsmmd.so"))

MAX_PARAMS = 8
NUM_SCALARS = 16

# enums (include/smmd.h)
F32, BF16 = 0, 1
K_DISTANCE, K_TANH_DISTANCE, K_DOT, K_RBF, K_MIX_RBF, K_MIX_RQ, K_TANH_MIX_RQ, K_POLY = range(8)
PREC_FP32, PREC_BF16, PREC_BF16X3, PREC_AUTO, PREC_FP16 = range(5)
EST_UNBIASED, EST_BIASED, EST_USTAT = range(3)
S_MMD2, S_SUM_XX, S_SUM_YY, S_SUM_XY, S_SUM_YX, S_DIAG_X, S_DIAG_Y, S_NONFINITE, S_VAR, S_RATIO = range(10)

PRECISIONS = {"fp32": PREC_FP32, "bf16": PREC_BF16, "bf16x3": PREC_BF16X3, "auto": PREC_AUTO, "fp16": PREC_FP16}
ESTIMATORS = {"unbiased": EST_UNBIASED, "biased": EST_BIASED, "u-statistic": EST_USTAT}

EXPORTS = [
    "smmd_version", "smmd_strerror", "smmd_last_cuda_error", "smmd_device_supported",
    "smmd_mmd2_workspace_bytes", "smmd_mmd2_fwd_bwd", "smmd_mmd2_fwd_bwd_gathered", "smmd_mmd2_combine",
    "smmd_peer_buffer_bytes", "smmd_mmd2_fwd_bwd_peers", "smmd_peer_set_pull_event", "smmd_draw_subsets_mt19937",
    "smmd_mmd2_and_ratio",
    "smmd_kernel_xy", "smmd_kernel_xy_bwd", "smmd_kernel_xy_bwd2", "smmd_kid_workspace_bytes", "smmd_kid_subsets",
    "smmd_poly_sums_workspace_bytes", "smmd_poly_sums", "smmd_kid_from_row_stats", "smmd_ratio_from_row_stats",
    "smmd_last_launch_count", "smmd_last_path", "smmd_profile_enable", "smmd_profile_last_ms", "smmd_profile_last_split_ms", "smmd_set_option",
]


class Problem(C.Structure):
    _fields_ = [
        ("m", C.c_int64), ("n", C.c_int64), ("d", C.c_int64),
        ("ldx", C.c_int64), ("ldy", C.c_int64),
        ("dtype", C.c_int32), ("kernel_id", C.c_int32), ("nparams", C.c_int32),
        ("params", C.c_float * MAX_PARAMS), ("wts", C.c_float * MAX_PARAMS),
        ("add_dot", C.c_float), ("degree", C.c_int32), ("biased", C.c_int32), ("precision", C.c_int32),
        ("rank", C.c_int32), ("world", C.c_int32),
    ]


MAX_PEERS = 16


class PeerTable(C.Structure):
    """smmd_peer_table (include/smmd.h): every rank's exchange buffer as mapped into this process."""
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("base", C.c_void_p * MAX_PEERS)]


class KidProblem(C.Structure):
    _fields_ = [
        ("n_g", C.c_int64), ("n_r", C.c_int64), ("d", C.c_int64),
        ("ldg", C.c_int64), ("ldr", C.c_int64),
        ("dtype", C.c_int32), ("n_subsets", C.c_int32), ("subset_size", C.c_int32), ("degree", C.c_int32),
        ("gamma", C.c_float), ("coef0", C.c_float),
        ("var_at_m", C.c_int64),
        ("mmd_est", C.c_int32), ("ret_var", C.c_int32), ("precision", C.c_int32),
        ("first_subset", C.c_int32), ("n_local", C.c_int32),
    ]


class SmmdError(RuntimeError):
    def __init__(self, status, where, detail=""):
        self.status = status
        super().__init__("%s failed: %s (status %d)%s" % (where, detail or "?", status, ""))


_lib = None


def load():
    """Load libsmmd.so (built in-tree by `make -C scaled-mmd-gan_b200/csrc` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            "libsmmd.so not found at %s -- build it with `make -C scaled-mmd-gan_b200/csrc` "
            "(there is no CPU/PyTorch fallback for this path)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i64, dbl = C.c_void_p, C.c_int64, C.c_double
    lib.smmd_version.restype = C.c_int
    lib.smmd_strerror.restype = C.c_char_p
    lib.smmd_strerror.argtypes = [C.c_int]
    lib.smmd_last_cuda_error.restype = C.c_char_p
    lib.smmd_device_supported.restype = C.c_int
    lib.smmd_last_launch_count.restype = C.c_int
    lib.smmd_last_path.restype = C.c_char_p
    lib.smmd_profile_enable.restype = None
    lib.smmd_profile_enable.argtypes = [C.c_int]
    lib.smmd_profile_last_ms.restype = C.c_float
    lib.smmd_mmd2_workspace_bytes.restype = C.c_size_t
    lib.smmd_mmd2_workspace_bytes.argtypes = [C.POINTER(Problem), C.c_int]
    lib.smmd_mmd2_fwd_bwd.restype = C.c_int
    lib.smmd_mmd2_fwd_bwd.argtypes = [C.POINTER(Problem), vp, vp, vp, vp, vp, vp, C.c_size_t, vp]
    lib.smmd_mmd2_fwd_bwd_gathered.restype = C.c_int
    lib.smmd_mmd2_fwd_bwd_gathered.argtypes = [C.POINTER(Problem), vp, i64, vp, vp, i64, vp, vp, vp, vp, C.c_size_t, vp]
    lib.smmd_draw_subsets_mt19937.restype = C.c_int
    lib.smmd_draw_subsets_mt19937.argtypes = [vp, C.POINTER(C.c_int32), i64, i64, C.c_int32, C.c_int32, vp, vp]
    lib.smmd_peer_set_pull_event.restype = C.c_int
    lib.smmd_peer_set_pull_event.argtypes = [vp]
    lib.smmd_peer_buffer_bytes.restype = C.c_size_t
    lib.smmd_peer_buffer_bytes.argtypes = [i64, i64]
    lib.smmd_mmd2_fwd_bwd_peers.restype = C.c_int
    lib.smmd_mmd2_fwd_bwd_peers.argtypes = [C.POINTER(Problem), C.POINTER(PeerTable), C.c_uint64, vp, vp, i64, vp, vp, vp, vp,
                                            C.c_size_t, vp]
    lib.smmd_mmd2_combine.restype = C.c_int
    lib.smmd_mmd2_combine.argtypes = [C.POINTER(Problem), vp, vp, vp]
    lib.smmd_mmd2_and_ratio.restype = C.c_int
    lib.smmd_mmd2_and_ratio.argtypes = [C.POINTER(Problem), vp, vp, dbl, vp, vp, C.c_size_t, vp]
    lib.smmd_kernel_xy.restype = C.c_int
    lib.smmd_kernel_xy.argtypes = [C.POINTER(Problem), vp, vp, vp, i64, vp]
    lib.smmd_kernel_xy_bwd.restype = C.c_int
    lib.smmd_kernel_xy_bwd.argtypes = [C.POINTER(Problem), vp, vp, vp, i64, vp, vp, vp]
    lib.smmd_kernel_xy_bwd2.restype = C.c_int
    lib.smmd_kernel_xy_bwd2.argtypes = [C.POINTER(Problem), vp, vp, vp, i64, vp, vp, vp, vp, vp, vp]
    lib.smmd_kid_workspace_bytes.restype = C.c_size_t
    lib.smmd_kid_workspace_bytes.argtypes = [C.POINTER(KidProblem)]
    lib.smmd_kid_subsets.restype = C.c_int
    lib.smmd_kid_subsets.argtypes = [C.POINTER(KidProblem), vp, vp, vp, vp, vp, vp, vp, C.c_size_t, vp]
    lib.smmd_poly_sums_workspace_bytes.restype = C.c_size_t
    lib.smmd_poly_sums_workspace_bytes.argtypes = [C.POINTER(KidProblem)]
    lib.smmd_poly_sums.restype = C.c_int
    lib.smmd_poly_sums.argtypes = [C.POINTER(KidProblem), vp, vp, vp, vp, C.c_size_t, vp]
    lib.smmd_profile_last_split_ms.restype = C.c_int
    lib.smmd_profile_last_split_ms.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.smmd_kid_from_row_stats.restype = C.c_int
    lib.smmd_kid_from_row_stats.argtypes = [vp, i64, C.c_int, C.c_int, i64, vp, vp, vp]
    lib.smmd_ratio_from_row_stats.restype = C.c_int
    lib.smmd_ratio_from_row_stats.argtypes = [vp, i64, C.c_int, C.c_int, dbl, dbl, vp, vp]
    lib.smmd_set_option.restype = C.c_int
    lib.smmd_set_option.argtypes = [C.c_char_p, C.c_longlong]
    _lib = lib
    return lib


def check(status, where):
    if status == 0:
        return
    lib = load()
    detail = lib.smmd_strerror(status).decode()
    if status == -6:
        detail += " [" + lib.smmd_last_cuda_error().decode() + "]"
    raise SmmdError(status, where, detail)


def last_path():
    return load().smmd_last_path().decode()


def last_launch_count():
    return int(load().smmd_last_launch_count())


options_epoch = 0   # bumped by set_option: cached workspace sizes depend on the selected path


def set_option(name, value):
    """Path-selection option of the library (include/smmd.h: smmd_set_option), e.g. set_option("sym_min_rows", 1)."""
    global options_epoch
    check(load().smmd_set_option(name.encode(), int(value)), "smmd_set_option(%s)" % name)
    options_epoch += 1


def profile_last_split_ms():
    """(first kernel ms, second kernel ms) of the last profiled two-kernel call, or None."""
    a, b = C.c_float(), C.c_float()
    if load().smmd_profile_last_split_ms(C.byref(a), C.byref(b)) != 0:
        return None
    return float(a.value), float(b.value)
