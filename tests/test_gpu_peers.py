"""Peer-memory entry point (smmd_mmd2_fwd_bwd_peers) at world = 1 on one GPU: the publish -> pull -> kernels -> sum
exchange -> combine chain through the C ABI must reproduce the plain call (same kernels on the same operand values), and
the one-launch small kernel must meet the exact tier against the fp64 oracle.  The multi-rank behaviour (flags crossing
GPUs, both slots, several steps) is covered by tests/multi_gpu_check.py under torchrun and by bench.py --gpus N."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import mmd_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _peers_call(spec, X, Y, precision, buf, step, biased=False):
    from smmd import _lib
    from smmd.mmd import _stream_ptr, _workspace

    lib = _lib.load()
    m, n, d = X.shape[0], Y.shape[0], X.shape[1]
    prob = spec.problem(m, n, d, d, d, torch.float32, biased, precision, 0, 1)
    table = _lib.PeerTable()
    table.world, table.rank = 1, 0
    table.base[0] = C.c_void_p(buf.data_ptr())
    nbytes = lib.smmd_mmd2_workspace_bytes(C.byref(prob), 1)
    ws = _workspace(nbytes, X.device)
    sc = torch.empty(_lib.NUM_SCALARS, dtype=torch.float64, device=X.device)
    dX, dY = torch.empty_like(X), torch.empty_like(Y)
    st = lib.smmd_mmd2_fwd_bwd_peers(C.byref(prob), C.byref(table), step, X.data_ptr(), Y.data_ptr(), d, sc.data_ptr(),
                                     dX.data_ptr(), dY.data_ptr(), ws.data_ptr(), nbytes, _stream_ptr(X.device))
    _lib.check(st, "smmd_mmd2_fwd_bwd_peers")
    return sc, dX, dY, _lib.last_path()


@pytest.mark.parametrize("shape,precision", [((700, 900, 128), "bf16"), ((640, 520, 512), "bf16"), ((1000, 1100, 192), "fp16")])
def test_peers_world1_matches_plain_call(shape, precision):
    from smmd import _lib, mmd

    m, n, d = shape
    rng = np.random.RandomState(m + d)
    X = torch.tensor((rng.randn(m, d) / np.sqrt(d)).astype(np.float32), device=DEV)
    Y = torch.tensor(((1.05 * rng.randn(n, d) + 0.1) / np.sqrt(d)).astype(np.float32), device=DEV)
    spec = mmd._mix_rq_kernel(X, Y).spec
    lib = _lib.load()
    buf = torch.zeros(int(lib.smmd_peer_buffer_bytes(m + n, d)), dtype=torch.uint8, device=DEV)
    ref_sc, ref_dX, ref_dY = mmd.fused_mmd2_raw(spec, X, Y, want_grad=True, precision=precision)
    ref_path = _lib.last_path()
    for step in (1, 2, 3):                  # both slots, flags that keep counting
        sc, dX, dY, path = _peers_call(spec, X, Y, precision, buf, step)
        assert path == ref_path
        assert sc[_lib.S_MMD2].item() == ref_sc[_lib.S_MMD2].item()
        assert torch.equal(dX, ref_dX) and torch.equal(dY, ref_dY)


@pytest.mark.parametrize("kernel,kw", [("mix_rq", {}), ("rbf", {}), ("mix_rbf", {"sigmas": [1.0, 2.0, 4.0, 8.0, 16.0]}),
                                       ("distance", {}), ("mix_rq_dot", {})])
@pytest.mark.parametrize("shape", [(64, 64, 1), (64, 64, 16), (100, 90, 33), (512, 512, 4), (300, 200, 64)])
def test_peers_one_launch_kernel_vs_oracle(kernel, kw, shape):
    from smmd import _lib, mmd

    m, n, d = shape
    rng = np.random.RandomState(7 * m + d)
    Xn = (rng.randn(m, d) / np.sqrt(d)).astype(np.float32)
    Yn = ((1.05 * rng.randn(n, d) + 0.1) / np.sqrt(d)).astype(np.float32)
    X, Y = torch.tensor(Xn, device=DEV), torch.tensor(Yn, device=DEV)
    spec = getattr(mmd, "_%s_kernel" % kernel)(X, Y, **kw).spec
    lib = _lib.load()
    buf = torch.zeros(int(lib.smmd_peer_buffer_bytes(m + n, d)), dtype=torch.uint8, device=DEV)
    for biased in (False, True):
        sc, dX, dY, path = _peers_call(spec, X, Y, "fp32", buf, 1, biased)
        assert path == "simt_fp32_small_peer"
        v, gx, gy = mmd_oracle.mmd2_and_grads(kernel, Xn, Yn, biased, np.float64, **kw)
        Kxx, Kxy, Kyy, _ = mmd_oracle.kernel_matrices(kernel, Xn, Yn, np.float64, **kw)
        kscale = max(abs(Kxx).mean(), abs(Kxy).mean(), abs(Kyy).mean())
        assert abs(sc[_lib.S_MMD2].item() - v) <= 1e-5 * abs(v) + 2e-7 * kscale
        assert np.abs(dX.cpu().numpy() - gx).max() <= 1e-5 * np.abs(gx).max()
        assert np.abs(dY.cpu().numpy() - gy).max() <= 1e-5 * np.abs(gy).max()


def test_peers_rejects_bad_tables_and_uncovered_shapes():
    from smmd import _lib, mmd

    X = torch.randn(2000, 16, device=DEV)
    Y = torch.randn(2000, 16, device=DEV)
    spec = mmd._mix_rq_kernel(X, Y).spec
    lib = _lib.load()
    buf = torch.zeros(int(lib.smmd_peer_buffer_bytes(4000, 16)), dtype=torch.uint8, device=DEV)
    with pytest.raises(_lib.SmmdError):           # exact tier beyond the one-launch kernel: no peer variant
        _peers_call(spec, X, Y, "fp32", buf, 1)


def test_peers_device_counted_steps():
    """step = 0: the kernels take (last completed step + 1) from the own buffer -- same results, counter advances, and the
    call can be captured into a CUDA graph and replayed."""
    from smmd import _lib, mmd

    rng = np.random.RandomState(5)
    X = torch.tensor(rng.randn(64, 16).astype(np.float32), device=DEV)
    Y = torch.tensor((1.1 * rng.randn(64, 16) + 0.1).astype(np.float32), device=DEV)
    spec = mmd._mix_rq_kernel(X, Y).spec
    lib = _lib.load()
    buf = torch.zeros(int(lib.smmd_peer_buffer_bytes(128, 16)), dtype=torch.uint8, device=DEV)
    counter = buf[512:520].view(torch.int64)
    ref = _peers_call(spec, X, Y, "fp32", buf, 1)            # host-given step 1
    assert counter.item() == 1
    for k in (2, 3, 4):
        sc, dX, dY, path = _peers_call(spec, X, Y, "fp32", buf, 0)
        assert path == "simt_fp32_small_peer" and counter.item() == k
        assert sc[_lib.S_MMD2].item() == ref[0][_lib.S_MMD2].item() and torch.equal(dX, ref[1]) and torch.equal(dY, ref[2])
    # tensor-core tier, device-counted
    Xb = torch.tensor((rng.randn(700, 128) / 11).astype(np.float32), device=DEV)
    Yb = torch.tensor((rng.randn(600, 128) / 11 + 0.01).astype(np.float32), device=DEV)
    specb = mmd._mix_rq_kernel(Xb, Yb).spec
    bufb = torch.zeros(int(lib.smmd_peer_buffer_bytes(1300, 128)), dtype=torch.uint8, device=DEV)
    r1 = _peers_call(specb, Xb, Yb, "bf16", bufb, 0)
    r2 = _peers_call(specb, Xb, Yb, "bf16", bufb, 0)
    assert bufb[512:520].view(torch.int64).item() == 2
    assert r1[0][_lib.S_MMD2].item() == r2[0][_lib.S_MMD2].item() and torch.equal(r1[1], r2[1])
    # CUDA-graph capture + replay of the one-launch loss
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        _peers_call(spec, X, Y, "fp32", buf, 0)              # warm the workspace cache of this stream
        torch.cuda.synchronize()
        before = counter.item()
        with torch.cuda.graph(g, stream=s):
            out = _peers_call(spec, X, Y, "fp32", buf, 0)
        for _ in range(5):
            g.replay()
    torch.cuda.synchronize()
    assert counter.item() == before + 5
    assert out[0][_lib.S_MMD2].item() == ref[0][_lib.S_MMD2].item() and torch.equal(out[1], ref[1])
