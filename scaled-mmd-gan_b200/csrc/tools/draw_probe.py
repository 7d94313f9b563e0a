"""Host timing of the subset draw (no GPU work, no torch): numpy's loop vs smmd_draw_subsets_mt19937 for several thread counts."""
import ctypes as C, os, sys, time
import numpy as np
lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "lib", "libsmmd.so"))
lib.smmd_draw_subsets_mt19937.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
def lib_draw(n, S, m):
    st = np.random.get_state(); key = np.ascontiguousarray(st[1], dtype=np.uint32).copy(); pos = C.c_int32(int(st[2]))
    ig = np.empty((S, m), np.int32); ir = np.empty((S, m), np.int32)
    assert lib.smmd_draw_subsets_mt19937(key.ctypes.data, C.byref(pos), n, n, S, m, ig.ctypes.data, ir.ctypes.data) == 0
    np.random.set_state((st[0], key, int(pos.value), st[3], st[4]))
    return ig, ir
n, S, m = 50000, 100, 1000
np.random.seed(0); t = time.perf_counter()
ref = [(np.random.choice(n, m, replace=False), np.random.choice(n, m, replace=False)) for _ in range(S)]
print("numpy loop: %.1f ms  (cpus %d)" % ((time.perf_counter() - t) * 1e3, os.cpu_count()))
for thr in ("1", "2", "4", "8", "16", None):
    if thr is None: os.environ.pop("SMMD_DRAW_THREADS", None)
    else: os.environ["SMMD_DRAW_THREADS"] = thr
    ts = []
    for _ in range(5):
        np.random.seed(0); t = time.perf_counter(); ig, ir = lib_draw(n, S, m); ts.append((time.perf_counter() - t) * 1e3)
    same = all(np.array_equal(ig[s], ref[s][0]) and np.array_equal(ir[s], ref[s][1]) for s in range(S))
    print("SMMD_DRAW_THREADS=%s: min %.1f ms, median %.1f ms  identical=%s" % (thr, min(ts), sorted(ts)[2], same))
