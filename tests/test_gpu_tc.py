"""GPU parity tests of the tcgen05 tensor-core path (bf16 operands, fp32 accumulate) against the fp64 numpy
oracle on the same seeded fp32 inputs, through the Python drop-in -> C ABI.

Tolerances (BASELINE.json north_star: rel 1e-3 for the bf16 tensor-core path, gradients element-wise):
  * MMD^2 / KID value: |v - v64| <= 1e-3 |v64| + floor, floor = 2e-6 * kscale (kscale = size of the block
    means that cancel in MMD^2; the floor is ~1/250 of bf16's unit roundoff 2^-9 on those means).
  * gradients, element-wise: |g - g64| <= 4e-3 * max|g64|.  W = a k'(D) and Z_j both enter the second
    UMMA rounded to bf16 (unit roundoff 2^-9 = 2e-3); measured 0.8e-3 .. 2.8e-3 of max|g| depending on
    how many columns average the rounding out.
"""
import numpy as np
import pytest
import torch

from oracle import kid_oracle, mmd_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _row_stacked_kernels():
    """This module pins the row-stacked kernels (fused tile-pair kernel, W row panels + GEMM): whole problems of >= 4096
    stacked rows would otherwise take the symmetric paths, which tests/test_gpu_sym.py covers."""
    from smmd import _lib

    _lib.set_option("sym", 0)
    yield
    _lib.set_option("sym", 1)
DEV = "cuda:0"

CASES = [
    ("rbf", {}),
    ("mix_rbf", {"sigmas": [1.0, 2.0, 4.0, 8.0, 16.0]}),          # ratio-4 gamma ladder variant
    ("mix_rbf", {}),                                                 # default sigmas -> generic variant
    ("mix_rq", {}),                                                  # default alphas -> 3-MUFU variant
    ("mix_rq", {"alphas": [0.2, 0.5, 1.0, 2.0, 5.0], "wts": [1.0, 0.5, 2.0, 1.0, 0.25]}),
    ("mix_rq_dot", {}),
    ("mix_rq_1dot", {}),
    ("tanh_mix_rq", {}),
    ("distance", {}),
    ("tanh_distance", {}),
]
SHAPES = [(300, 200, 100), (1000, 1000, 128), (513, 700, 256), (129, 127, 64), (2048, 2048, 192)]


def _data(m, n, d, seed):
    rng = np.random.RandomState(seed)
    X = (rng.randn(m, d) / np.sqrt(d)).astype(np.float32)
    Y = ((1.05 * rng.randn(n, d) + 0.1) / np.sqrt(d)).astype(np.float32)
    return X, Y


def _fused_path(d):
    """d <= 256: single fused kernel (O in tensor memory); above: W row panels + GEMM."""
    return "tc_bf16_fused" if d <= 256 else "tc_bf16_wz"


def _kscale(name, kw, X, Y):
    n = min(len(X), 256)
    Kxx, Kxy, Kyy, _ = mmd_oracle.kernel_matrices(name, X[:n], Y[:n], np.float64, **kw)
    return max(abs(Kxx).mean(), abs(Kxy).mean(), abs(Kyy).mean())


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("case", CASES, ids=lambda c: c[0] + ("+" if c[1] else ""))
def test_tc_fused_fwd_bwd_vs_oracle(case, shape):
    from smmd import _lib, mmd

    name, kw = case
    m, n, d = shape
    X, Y = _data(m, n, d, m + n + d)
    for biased in (False, True):
        Xt = torch.tensor(X, device=DEV, requires_grad=True)
        Yt = torch.tensor(Y, device=DEV, requires_grad=True)
        loss = mmd.mmd2(getattr(mmd, "_%s_kernel" % name)(Xt, Yt, **kw), biased=biased, precision="bf16")
        loss.backward()
        assert _lib.last_path() == _fused_path(d)
        v, gx, gy = mmd_oracle.mmd2_and_grads(name, X, Y, biased, np.float64, **kw)
        floor = 2e-6 * _kscale(name, kw, X, Y)
        assert abs(loss.item() - v) <= 1e-3 * abs(v) + floor, (name, biased, loss.item(), v)
        for got, ref in ((Xt.grad, gx), (Yt.grad, gy)):
            err = np.abs(got.cpu().numpy().astype(np.float64) - ref).max()
            assert err <= 4e-3 * np.abs(ref).max(), (name, biased, err, np.abs(ref).max())


# BASELINE config 4 sweeps d = 256..1024: the two-pass path (bf16 W row panels, then O = W Z as a GEMM) against the
# fp64 oracle, ragged sizes included
WIDE_SHAPES = [(700, 900, 512), (513, 640, 384), (300, 200, 300), (640, 520, 1024), (1100, 1000, 768), (257, 255, 600),
               (384, 400, 2048)]
WIDE_CASES = [("mix_rq", {}), ("rbf", {}), ("mix_rbf", {"sigmas": [1.0, 2.0, 4.0, 8.0, 16.0]}), ("mix_rq_dot", {}),
              ("tanh_mix_rq", {}), ("distance", {})]


@pytest.mark.parametrize("shape", WIDE_SHAPES, ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("case", WIDE_CASES, ids=lambda c: c[0] + ("+" if c[1] else ""))
def test_tc_wide_fwd_bwd_vs_oracle(case, shape):
    from smmd import _lib, mmd

    name, kw = case
    m, n, d = shape
    X, Y = _data(m, n, d, m + n + d)
    for biased in (False, True):
        Xt = torch.tensor(X, device=DEV, requires_grad=True)
        Yt = torch.tensor(Y, device=DEV, requires_grad=True)
        loss = mmd.mmd2(getattr(mmd, "_%s_kernel" % name)(Xt, Yt, **kw), biased=biased, precision="bf16")
        loss.backward()
        assert _lib.last_path() == _fused_path(d)
        v, gx, gy = mmd_oracle.mmd2_and_grads(name, X, Y, biased, np.float64, **kw)
        floor = 2e-6 * _kscale(name, kw, X, Y)
        assert abs(loss.item() - v) <= 1e-3 * abs(v) + floor, (name, biased, loss.item(), v)   # bf16 TC tolerance
        for got, ref in ((Xt.grad, gx), (Yt.grad, gy)):
            err = np.abs(got.cpu().numpy().astype(np.float64) - ref).max()
            assert err <= 4e-3 * np.abs(ref).max(), (name, biased, err, np.abs(ref).max())


def test_tc_wide_long_stream_and_shards():
    """Many tiles per CTA (ring phases wrap many times), determinism, and rank/world row shards at d = 512."""
    from smmd import _lib, mmd

    X, Y = _data(5000, 4600, 512, 77)
    Xt, Yt = torch.tensor(X, device=DEV), torch.tensor(Y, device=DEV)
    spec = mmd._mix_rq_kernel(Xt, Yt).spec
    _lib.set_option("sym", 0)   # whole problems of this size take the symmetric path by default (tests/test_gpu_sym.py)
    try:
        full, gX, gY = mmd.fused_mmd2_raw(spec, Xt, Yt, precision="bf16")
        assert _lib.last_path() == "tc_bf16_wz"
        full2, gX2, gY2 = mmd.fused_mmd2_raw(spec, Xt, Yt, precision="bf16")
    finally:
        _lib.set_option("sym", 0)   # (module default, see _row_stacked_kernels)
    ref, rX, rY = mmd.fused_mmd2_raw(spec, Xt, Yt, precision="fp32")
    assert abs(full[_lib.S_MMD2].item() - ref[_lib.S_MMD2].item()) <= 1e-3 * abs(ref[_lib.S_MMD2].item())
    assert (gX - rX).abs().max() <= 4e-3 * rX.abs().max()
    assert (gY - rY).abs().max() <= 4e-3 * rY.abs().max()
    assert torch.equal(gX, gX2) and torch.equal(gY, gY2) and torch.equal(full, full2)   # deterministic
    world = 3
    acc = torch.zeros_like(full)
    for rank in range(world):
        sc, dX, dY = mmd.fused_mmd2_raw(spec, Xt, Yt, precision="bf16", rank=rank, world=world)
        acc += sc
        x0, x1 = 5000 * rank // world, 5000 * (rank + 1) // world
        y0, y1 = 4600 * rank // world, 4600 * (rank + 1) // world
        # shards cut the tile stream and the GEMM's K range at different places: fp32 accumulation order of the
        # row sums r_i and of O_i differs, and g_i = r_i z_i - O_i cancels ~10x (same bound as multi_gpu_check.py)
        ex, ey = (dX - gX[x0:x1]).abs().max().item(), (dY - gY[y0:y1]).abs().max().item()
        assert ex <= 2e-4 * gX.abs().max().item(), (rank, ex, gX.abs().max().item())
        assert ey <= 2e-4 * gY.abs().max().item(), (rank, ey, gY.abs().max().item())
    for i in (_lib.S_SUM_XX, _lib.S_SUM_YY, _lib.S_SUM_XY, _lib.S_SUM_YX):
        assert abs(acc[i].item() - full[i].item()) <= 1e-9 * abs(full[i].item())


def test_tc_wide_cta_pair_path():
    """>= 148 row blocks per panel: pass 1 runs as CTA pairs (cta_group::2 UMMAs, each CTA supplies half of every
    Z_j tile).  Odd number of row blocks (dummy second block in the last pair), m != n, checked against the exact
    fp32 path on the GPU (itself pinned to the oracle) with the bf16 tolerance."""
    from smmd import _lib, mmd

    m, n, d = 9700, 9500, 320          # 76 + 75 = 151 row blocks
    X, Y = _data(m, n, d, 5)
    Xt, Yt = torch.tensor(X, device=DEV), torch.tensor(Y, device=DEV)
    spec = mmd._mix_rq_kernel(Xt, Yt).spec
    _lib.set_option("sym", 0)   # (whole problems of this size take the symmetric path by default; row shards take this one)
    try:
        sc, gX, gY = mmd.fused_mmd2_raw(spec, Xt, Yt, precision="bf16")
        assert _lib.last_path() == "tc_bf16_wz_pair"
        sc2, gX2, gY2 = mmd.fused_mmd2_raw(spec, Xt, Yt, precision="bf16")
    finally:
        _lib.set_option("sym", 0)   # (module default, see _row_stacked_kernels)
    rs, rX, rY = mmd.fused_mmd2_raw(spec, Xt, Yt, precision="fp32")
    assert abs(sc[_lib.S_MMD2].item() - rs[_lib.S_MMD2].item()) <= 1e-3 * abs(rs[_lib.S_MMD2].item())
    for i in (_lib.S_SUM_XX, _lib.S_SUM_YY, _lib.S_SUM_XY, _lib.S_SUM_YX):
        assert abs(sc[i].item() - rs[i].item()) <= 1e-4 * abs(rs[i].item())
    assert (gX - rX).abs().max() <= 4e-3 * rX.abs().max()
    assert (gY - rY).abs().max() <= 4e-3 * rY.abs().max()
    assert torch.equal(gX, gX2) and torch.equal(gY, gY2) and torch.equal(sc, sc2)   # deterministic


def test_tc_wide_bf16_inputs_and_strided_rows():
    """Two-pass path with bf16 critic features and with row-strided fp32 views (ld > d), against the fp64 oracle on
    exactly the values the kernel sees."""
    from smmd import _lib, mmd

    m, n, d = 520, 600, 448
    X, Y = _data(m, n, d, 11)
    # (a) bf16 inputs: oracle on the bf16-rounded values
    Xb = torch.tensor(X, device=DEV).to(torch.bfloat16).requires_grad_(True)
    Yb = torch.tensor(Y, device=DEV).to(torch.bfloat16).requires_grad_(True)
    loss = mmd.mmd2(mmd._mix_rq_kernel(Xb, Yb), precision="bf16")
    loss.backward()
    assert _lib.last_path() == "tc_bf16_wz"
    Xr, Yr = Xb.detach().float().cpu().numpy(), Yb.detach().float().cpu().numpy()
    v, gx, gy = mmd_oracle.mmd2_and_grads("mix_rq", Xr, Yr, False, np.float64)
    assert abs(loss.item() - v) <= 1e-3 * abs(v) + 2e-6 * _kscale("mix_rq", {}, Xr, Yr)
    for got, ref in ((Xb.grad, gx), (Yb.grad, gy)):
        assert got.dtype == torch.bfloat16
        err = np.abs(got.float().cpu().numpy().astype(np.float64) - ref).max()
        assert err <= 8e-3 * np.abs(ref).max(), (err, np.abs(ref).max())   # + bf16 rounding of the returned gradient
    # (b) strided rows: a [m, d] view into a wider buffer
    big = torch.zeros((m, d + 72), device=DEV)
    big[:, :d] = torch.tensor(X, device=DEV)
    Xs = big[:, :d].requires_grad_(True)
    Yt = torch.tensor(Y, device=DEV, requires_grad=True)
    loss = mmd.mmd2(mmd._mix_rq_kernel(Xs, Yt), precision="bf16")
    gxs, gys = torch.autograd.grad(loss, [Xs, Yt])
    v, gx, gy = mmd_oracle.mmd2_and_grads("mix_rq", X, Y, False, np.float64)
    assert abs(loss.item() - v) <= 1e-3 * abs(v) + 2e-6 * _kscale("mix_rq", {}, X, Y)
    assert np.abs(gxs.cpu().numpy() - gx).max() <= 4e-3 * np.abs(gx).max()
    assert np.abs(gys.cpu().numpy() - gy).max() <= 4e-3 * np.abs(gy).max()


_PANEL_SNIPPET = r"""
import sys, numpy as np, torch
sys.path.insert(0, {pkg!r})
from smmd import _lib, mmd
rng = np.random.RandomState(3)
X = torch.tensor((rng.randn(1500, 640) / 25.3).astype(np.float32), device="cuda")
Y = torch.tensor(((1.05 * rng.randn(1300, 640) + 0.1) / 25.3).astype(np.float32), device="cuda")
spec = mmd._mix_rq_kernel(X, Y).spec
full, gX, gY = mmd.fused_mmd2_raw(spec, X, Y, precision="bf16")
assert _lib.last_path() == "tc_bf16_wz", _lib.last_path()
torch.save({{"sc": full.cpu(), "gX": gX.cpu(), "gY": gY.cpu()}}, sys.argv[1])
"""


def test_tc_wide_panel_size_does_not_change_results(tmp_path):
    """The W row-panel budget only changes how the work is cut: a 1 MB budget (11 panels of 2 row blocks here)
    must give the same gradients as one panel (same tiles, same K order; only the GEMM's K split may differ)."""
    import os
    import subprocess
    import sys

    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scaled-mmd-gan_b200")
    outs = []
    for mb in ("1", "4096"):
        out = tmp_path / ("panel_%s.pt" % mb)
        env = dict(os.environ, SMMD_WZ_PANEL_MB=mb)
        subprocess.run([sys.executable, "-c", _PANEL_SNIPPET.format(pkg=pkg), str(out)], check=True, env=env, timeout=300)
        outs.append(torch.load(out))
    a, b = outs
    ex, ey = (a["gX"] - b["gX"]).abs().max().item(), (a["gY"] - b["gY"]).abs().max().item()
    print("panel-size gradient deviation (relative to max|g|):", ex / b["gX"].abs().max().item(), ey / b["gY"].abs().max().item())
    assert ex <= 2e-4 * b["gX"].abs().max().item() and ey <= 2e-4 * b["gY"].abs().max().item()
    assert abs(a["sc"][0].item() - b["sc"][0].item()) <= 1e-9 * abs(b["sc"][0].item())


@pytest.mark.parametrize("name,kw", [("rbf", {}), ("mix_rq", {}), ("distance", {}),
                                      ("mix_rbf", {"sigmas": [1.0, 2.0, 4.0, 8.0, 16.0]})])
@pytest.mark.parametrize("precision", ["bf16", "bf16x3"])
def test_tc_value_only_stream_path(name, kw, precision):
    from smmd import _lib, mmd

    X, Y = _data(900, 1100, 320, 5)   # d > 256: value-only path streams K, any d
    Xt, Yt = torch.tensor(X, device=DEV), torch.tensor(Y, device=DEV)
    got = mmd.mmd2(getattr(mmd, "_%s_kernel" % name)(Xt, Yt, **kw), precision=precision).item()
    assert _lib.last_path().startswith("tc_bf16")
    v = mmd_oracle.mmd2(name, X, Y, False, np.float64, **kw)
    tol = 1e-3 if precision == "bf16" else 1e-4
    assert abs(got - v) <= tol * abs(v) + 2e-6 * _kscale(name, kw, X, Y), (got, v)


def test_fused_kernel_is_repeatable_under_stress():
    """A few hundred back-to-back launches of the fused kernel (tile-pair variant) on shapes with many tiles per CTA,
    several row-block units per CTA and special (diagonal / padded) tiles: every launch must return bit-identical
    sums and gradients (fixed reduction order), and none may trip the bounded barrier waits (a lost or doubled
    mbarrier phase shows up as a CUDA error here, not as a hang)."""
    from smmd import _lib, mmd

    for (m, n, d, reps) in ((3000, 5000, 192, 150), (4100, 4000, 256, 150), (1000, 1100, 128, 200)):
        X, Y = _data(m, n, d, 3)
        Xt, Yt = torch.tensor(X, device=DEV), torch.tensor(Y, device=DEV)
        spec = mmd._mix_rq_kernel(Xt, Yt).spec
        ref, rX, rY = mmd.fused_mmd2_raw(spec, Xt, Yt, precision="bf16")
        assert _lib.last_path() == "tc_bf16_fused"
        for _ in range(reps):
            sc, gX, gY = mmd.fused_mmd2_raw(spec, Xt, Yt, precision="bf16")
        torch.cuda.synchronize()
        assert torch.equal(sc, ref) and torch.equal(gX, rX) and torch.equal(gY, rY)


def test_auto_precision_dispatch_and_refusals():
    from smmd import _lib, mmd

    X, Y = _data(64, 64, 16, 1)
    Xt, Yt = torch.tensor(X, device=DEV, requires_grad=True), torch.tensor(Y, device=DEV, requires_grad=True)
    mmd.mmd2(mmd._mix_rq_kernel(Xt, Yt)).backward()                 # small -> exact path, single launch
    assert _lib.last_path() == "simt_fp32_small"
    Xb, Yb = _data(1024, 1024, 128, 2)
    Xt, Yt = torch.tensor(Xb, device=DEV, requires_grad=True), torch.tensor(Yb, device=DEV, requires_grad=True)
    mmd.mmd2(mmd._mix_rq_kernel(Xt, Yt)).backward()                 # large -> tensor cores
    assert _lib.last_path() == "tc_bf16_fused"
    mmd.mmd2(mmd._dot_kernel(Xt, Yt)).backward()                    # dot is closed-form territory -> exact path
    assert _lib.last_path() == "simt_fp32"
    with pytest.raises(_lib.SmmdError):                             # explicit request that cannot be honoured
        mmd.mmd2(mmd._dot_kernel(Xt, Yt), precision="bf16")
    Xw, Yw = _data(512, 512, 300, 3)
    Xt, Yt = torch.tensor(Xw, device=DEV, requires_grad=True), torch.tensor(Yw, device=DEV, requires_grad=True)
    mmd.mmd2(mmd._mix_rq_kernel(Xt, Yt)).backward()                 # AUTO, d > 256 -> two-pass tensor-core path
    assert _lib.last_path() == "tc_bf16_wz"


def test_row_shards_compose_to_single_gpu_result():
    """smmd_problem.rank/world on ONE GPU: partial sums of all shards add up to the unsharded sums and each
    shard's gradient rows equal the corresponding rows of the unsharded gradient."""
    from smmd import _lib, mmd

    X, Y = _data(1000, 1200, 128, 9)
    Xt, Yt = torch.tensor(X, device=DEV), torch.tensor(Y, device=DEV)
    for precision in ("fp32", "bf16"):
        spec = mmd._mix_rq_kernel(Xt, Yt).spec
        full, gX, gY = mmd.fused_mmd2_raw(spec, Xt, Yt, precision=precision)
        world = 4
        acc = torch.zeros_like(full)
        for rank in range(world):
            sc, dX, dY = mmd.fused_mmd2_raw(spec, Xt, Yt, precision=precision, rank=rank, world=world)
            acc += sc
            x0, x1 = 1000 * rank // world, 1000 * (rank + 1) // world
            y0, y1 = 1200 * rank // world, 1200 * (rank + 1) // world
            # shards cut the tile stream at different places -> fp32 accumulation order differs slightly
            assert (dX - gX[x0:x1]).abs().max() <= 1e-5 * gX.abs().max()
            assert (dY - gY[y0:y1]).abs().max() <= 1e-5 * gY.abs().max()
        for i in (_lib.S_SUM_XX, _lib.S_SUM_YY, _lib.S_SUM_XY, _lib.S_SUM_YX, _lib.S_DIAG_X, _lib.S_DIAG_Y):
            assert abs(acc[i].item() - full[i].item()) <= 1e-9 * abs(full[i].item())
        assert abs(full[_lib.S_SUM_XY].item() - full[_lib.S_SUM_YX].item()) <= 1e-6 * abs(full[_lib.S_SUM_XY].item())


@pytest.mark.parametrize("precision,tol", [("bf16x3", 1e-4), ("bf16", 2e-2)])
def test_tc_kid_vs_oracle(precision, tol):
    """KID on tensor cores.  The split-bf16 (3-term) Gram is the default: cubing amplifies input rounding
    x3 and KID itself is a ~1e-3 difference of O(1) block means, so plain bf16 operands only reach ~1e-2
    on KID (kept selectable, reported honestly); bf16x3 is at the fp32 level."""
    from smmd import _lib, compute_scores

    g = np.maximum(np.random.RandomState(1234).randn(3000, 512), 0).astype(np.float32)
    r = np.maximum(np.random.RandomState(1235).randn(3000, 512) + 0.02, 0).astype(np.float32)
    np.random.seed(0)
    mm, vv = compute_scores.polynomial_mmd_averages(g, r, n_subsets=6, subset_size=500, ret_var=True,
                                                    precision=precision)
    assert _lib.last_path() == ("tc_bf16x3_kid" if precision == "bf16x3" else "tc_bf16_kid")
    np.random.seed(0)
    rm, rv = kid_oracle.polynomial_mmd_averages(g.astype(np.float64), r.astype(np.float64), n_subsets=6,
                                                subset_size=500, ret_var=True)
    assert np.abs(mm - rm).max() <= tol * np.abs(rm).max(), (mm, rm)
    assert np.abs(vv - rv).max() <= max(50 * tol, 2e-2) * np.abs(rv).max(), (vv, rv)
    # default precision for KID is the split path
    np.random.seed(0)
    m2 = compute_scores.polynomial_mmd_averages(g, r, n_subsets=6, subset_size=500, ret_var=False)
    assert _lib.last_path() == "tc_bf16x3_kid_sym"   # totals only -> symmetric upper-triangle enumeration
    assert np.abs(m2 - rm).max() <= 1e-4 * np.abs(rm).max()


def test_tc_kid_properties_at_scale():
    """Size-independent properties at a larger size than the oracle is run at: permuting the rows inside a
    subset leaves its KID unchanged; identical subsets give identical values; swapping g and r swaps nothing
    (the unbiased estimator is symmetric)."""
    from smmd import compute_scores

    dev = torch.device(DEV)
    gen = torch.Generator(device=dev).manual_seed(1)
    g = torch.relu(torch.randn(8000, 2048, device=dev, generator=gen))
    r = torch.relu(torch.randn(8000, 2048, device=dev, generator=gen) + 0.02)
    idx = torch.stack([torch.randperm(8000, device=dev, generator=gen)[:1000] for _ in range(4)]).to(torch.int32)
    idx2 = idx.clone()
    idx2[1] = idx[0][torch.randperm(1000, device=dev, generator=gen)]   # subset 1 := permutation of subset 0
    a, _ = compute_scores.kid_subsets(g, r, idx, idx, ret_var=False)
    b, _ = compute_scores.kid_subsets(g, r, idx2, idx2, ret_var=False)
    # same multiset of rows in another order: only the fp32 accumulation order inside the tiles changes
    # (KID is a ~5e-4 difference of O(1) block means, so 1e-7-level reordering noise shows up at ~1e-4 relative)
    assert abs(b[1].item() - b[0].item()) <= 1e-9 + 1e-3 * abs(b[0].item())
    assert torch.allclose(a[[0, 2, 3]], b[[0, 2, 3]], rtol=1e-9, atol=1e-12)
    c, _ = compute_scores.kid_subsets(r, g, idx, idx, ret_var=False)
    # swapping the roles reorders the three split-bf16 partial products inside the fp32 accumulation
    assert torch.allclose(a, c, rtol=1e-3, atol=1e-8), (a, c)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_matches_single_gpu():
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29631",
                          os.path.join(root, "tests", "multi_gpu_check.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "MULTI_GPU_OK" in out.stdout
