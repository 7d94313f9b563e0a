"""CPU (gloo, world_size 2): host-side logic of the row-sharded MMD^2 and subset-sharded KID -- gather
order, shard arithmetic, partial-sum all-reduce and the combine formula -- with the per-rank device
compute replaced by an oracle-backed emulation of the C ABI's shard contract (include/smmd.h:
smmd_problem.rank/world, scalars[] layout).  The sharded result must equal the single-device oracle on
the concatenated global batch."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import kid_oracle, mmd_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _emulated_local_compute(spec, gathered, Xl, Yl, m, n, biased, precision, rank, world):
    """What smmd_mmd2_fwd_bwd_gathered returns for (rank, world), computed with the fp64 oracle from the
    block-interleaved gathered layout (world blocks of [X_local ; Y_local])."""
    from smmd import _lib
    from smmd.distributed import shard_rows

    G = gathered.numpy().astype(np.float64)
    ml, nl = m // world, n // world
    blocks = G.reshape(world, ml + nl, -1)
    X = blocks[:, :ml].reshape(m, -1)
    Y = blocks[:, ml:].reshape(n, -1)
    x0, x1 = shard_rows(m, rank, world)
    y0, y1 = shard_rows(n, rank, world)
    kw = {"alphas": spec.params, "wts": spec.wts, "add_dot": spec.add_dot}
    Kxx, Kxy, Kyy, cd = mmd_oracle.kernel_matrices("mix_rq", X, Y, np.float64, **kw)
    sc = np.zeros(_lib.NUM_SCALARS)
    off = lambda K, lo, hi: K[lo:hi].sum() - np.trace(K[lo:hi, lo:hi])
    sc[_lib.S_SUM_XX] = off(Kxx, x0, x1)
    sc[_lib.S_SUM_YY] = off(Kyy, y0, y1)
    sc[_lib.S_SUM_XY] = Kxy[x0:x1].sum()
    sc[_lib.S_SUM_YX] = Kxy[:, y0:y1].sum()
    sc[_lib.S_DIAG_X] = np.diagonal(Kxx)[x0:x1].sum()
    sc[_lib.S_DIAG_Y] = np.diagonal(Kyy)[y0:y1].sum()
    _, gX, gY = mmd_oracle.mmd2_and_grads("mix_rq", X, Y, biased, np.float64, **kw)
    return (torch.tensor(sc), torch.tensor(gX[x0:x1], dtype=torch.float32), torch.tensor(gY[y0:y1], dtype=torch.float32))


def _emulated_combine(spec, sums, m, n, d, biased, dtype):
    """Same arithmetic as mmd2_from_sums in csrc/smmd_simt.cu."""
    from smmd import _lib

    s = sums.numpy()
    cross = s[_lib.S_SUM_XY] + s[_lib.S_SUM_YX]
    if biased:
        v = (s[_lib.S_SUM_XX] + s[_lib.S_DIAG_X]) / (m * m) + (s[_lib.S_SUM_YY] + s[_lib.S_DIAG_Y]) / (n * n) - cross / (m * n)
    else:
        cd = float(spec.const_diagonal)
        v = ((s[_lib.S_SUM_XX] + s[_lib.S_DIAG_X] - m * cd) / (m * (m - 1))
             + (s[_lib.S_SUM_YY] + s[_lib.S_DIAG_Y] - n * cd) / (n * (n - 1)) - cross / (m * n))
    return torch.tensor(v, dtype=torch.float64)


def _emulated_kid(codes_g, codes_r, idx_g, idx_r, var_at_m=None, ret_var=True, precision=None, first_subset=0,
                  n_local=0, **kw):
    S = idx_g.shape[0]
    mm = torch.zeros(S, dtype=torch.float64)
    vv = torch.zeros(S, dtype=torch.float64)
    g, r = codes_g.numpy().astype(np.float64), codes_r.numpy().astype(np.float64)
    for s in range(first_subset, first_subset + n_local):
        out = kid_oracle.polynomial_mmd(g[idx_g[s].numpy()], r[idx_r[s].numpy()], var_at_m=var_at_m, ret_var=True)
        mm[s], vv[s] = out
    return mm, (vv if ret_var else None)


def _worker(rank, world, port, biased, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "scaled-mmd-gan_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from smmd import _lib, mmd
        from smmd.distributed import sharded_mmd2, sharded_polynomial_mmd_averages

        b, d = 12, 5
        rng = np.random.RandomState(100 + rank)
        Xl = torch.tensor(rng.randn(b, d).astype(np.float32), requires_grad=True)
        Yl = torch.tensor((1.1 * rng.randn(b, d) + 0.1).astype(np.float32), requires_grad=True)
        # build the handle without the CUDA-only checks (host-logic test)
        spec = mmd.KernelSpec(_lib.K_MIX_RQ, [0.1, 1.0, 10.0], [1.0, 1.0, 1.0], add_dot=0.1, const_diagonal=3.0)
        K = mmd.KernelHandle.__new__(mmd.KernelHandle)
        K.spec, K.X, K.Y, K._dense = spec, Xl, Yl, None
        loss = sharded_mmd2(K, biased=biased, _local_compute=_emulated_local_compute, _combine=_emulated_combine)
        (loss * 2.0).backward()
        # KID: subsets split across ranks
        g = torch.tensor(np.maximum(np.random.RandomState(1).randn(60, 16), 0).astype(np.float32))
        r = torch.tensor(np.maximum(np.random.RandomState(2).randn(70, 16) + 0.02, 0).astype(np.float32))
        ig, ir = kid_oracle.draw_subset_indices(60, 70, 5, 20, np.random.RandomState(0))
        mm, vv = sharded_polynomial_mmd_averages(g, r, torch.tensor(ig), torch.tensor(ir), ret_var=True,
                                                 _local_kid=_emulated_kid)
        q.put((rank, float(loss), Xl.grad.numpy(), Yl.grad.numpy(), Xl.detach().numpy(), Yl.detach().numpy(),
               mm.numpy(), vv.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("biased", [False, True])
def test_sharded_mmd2_and_kid_world2(biased):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + (1 if biased else 0)
    procs = [ctx.Process(target=_worker, args=(r, world, port, biased, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    X = np.concatenate([r[4] for r in res])
    Y = np.concatenate([r[5] for r in res])
    v, gX, gY = mmd_oracle.mmd2_and_grads("mix_rq_dot", X, Y, biased, np.float64)
    for rank, loss, gx, gy, *_ in res:
        assert abs(loss - v) <= 1e-6 * abs(v)                       # same scalar on every rank
        assert np.allclose(gx, 2.0 * gX[rank * 12:(rank + 1) * 12], rtol=1e-5, atol=1e-9)
        assert np.allclose(gy, 2.0 * gY[rank * 12:(rank + 1) * 12], rtol=1e-5, atol=1e-9)
    g = np.maximum(np.random.RandomState(1).randn(60, 16), 0).astype(np.float32).astype(np.float64)
    r_ = np.maximum(np.random.RandomState(2).randn(70, 16) + 0.02, 0).astype(np.float32).astype(np.float64)
    ig, ir = kid_oracle.draw_subset_indices(60, 70, 5, 20, np.random.RandomState(0))
    ref_m, ref_v = kid_oracle.polynomial_mmd_averages(g, r_, n_subsets=5, subset_size=20, ret_var=True, idx_g=ig, idx_r=ir)
    for rr in res:
        assert np.allclose(rr[6], ref_m, rtol=1e-10) and np.allclose(rr[7], ref_v, rtol=1e-10)


def test_shard_arithmetic_matches_c_abi_contract():
    from smmd.distributed import kid_shard, shard_rows

    for total in (1, 7, 64, 100, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard_rows(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert sum(kid_shard(total, r, world)[1] for r in range(world)) == total


def _mismatch_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "scaled-mmd-gan_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from smmd.distributed import check_equal_shards

        try:
            check_equal_shards(12 + rank, 12, torch.device("cpu"))    # rank 1 holds one more fake row
            q.put((rank, "no error"))
        except ValueError as e:
            q.put((rank, str(e)))
        check_equal_shards(12, 9, torch.device("cpu"))                # equal on every rank: passes
    finally:
        dist.destroy_process_group()


def test_unequal_shards_raise_on_every_rank():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_mismatch_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all("different numbers of rows" in msg for _, msg in res), res


def _emulated_peer_compute(spec, px, Xl, Yl, m, n, biased, precision):
    """What smmd_mmd2_fwd_bwd_peers leaves on every rank (COMBINED scalars + local gradients), with the kernels' exchange
    over peer memory emulated by gloo collectives and the arithmetic by the fp64 oracle."""
    from smmd import _lib

    world, rank = px.world, px.rank
    step = px.next_step()
    Xs = [torch.empty_like(Xl) for _ in range(world)]
    Ys = [torch.empty_like(Yl) for _ in range(world)]
    dist.all_gather(Xs, Xl.detach())
    dist.all_gather(Ys, Yl.detach())
    X, Y = torch.cat(Xs).numpy().astype(np.float64), torch.cat(Ys).numpy().astype(np.float64)
    v, gX, gY = mmd_oracle.mmd2_and_grads("mix_rq_dot", X, Y, biased, np.float64)
    ml, nl = m // world, n // world
    sc = torch.zeros(_lib.NUM_SCALARS, dtype=torch.float64)
    sc[_lib.S_MMD2] = v
    sc[_lib.S_VAR] = float(step)       # (smuggles the step number out for the sequencing assertion)
    return (sc, torch.tensor(gX[rank * ml:(rank + 1) * ml].astype(np.float32)),
            torch.tensor(gY[rank * nl:(rank + 1) * nl].astype(np.float32)))


def _peer_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "scaled-mmd-gan_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from smmd import _lib, mmd
        from smmd.distributed import PeerExchange, sharded_mmd2

        def fake_map(nbytes, device, group):     # every rank "maps" buffers at made-up, 256-aligned addresses
            own = torch.zeros(nbytes, dtype=torch.uint8)
            mine = torch.tensor([0x10000000 * (rank + 1)], dtype=torch.int64)
            ptrs = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(ptrs, mine)
            return own, [int(p) for p in ptrs], None

        b, d = 12, 5
        px = PeerExchange(2 * b, d, "cpu", map_buffers=fake_map, _compute=_emulated_peer_compute)
        table = (px.table.world, px.table.rank, [int(px.table.base[r] or 0) for r in range(world)])
        assert px.fits(2 * b, d) and not px.fits(2 * b + 1, d + 60)
        out = []
        for it in range(3):
            rng = np.random.RandomState(100 + 7 * it + rank)
            Xl = torch.tensor(rng.randn(b, d).astype(np.float32), requires_grad=True)
            Yl = torch.tensor((1.1 * rng.randn(b, d) + 0.1).astype(np.float32), requires_grad=True)
            spec = mmd.KernelSpec(_lib.K_MIX_RQ, [0.1, 1.0, 10.0], [1.0, 1.0, 1.0], add_dot=0.1, const_diagonal=3.0)
            K = mmd.KernelHandle.__new__(mmd.KernelHandle)
            K.spec, K.X, K.Y, K._dense = spec, Xl, Yl, None
            loss = sharded_mmd2(K, exchange=px)
            (loss * 3.0).backward()
            out.append((float(loss), Xl.grad.numpy(), Yl.grad.numpy(), Xl.detach().numpy(), Yl.detach().numpy()))
        q.put((rank, table, px.step, px.nbytes, out))
    finally:
        dist.destroy_process_group()


def test_peer_exchange_host_logic_world2():
    """PeerExchange: table construction from the mapped addresses, buffer size from the library, one step number per
    call (the same on every rank), and the autograd plumbing of sharded_mmd2(..., exchange=px)."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_peer_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, table, step, nbytes, out in res:
        assert table == (world, rank, [0x10000000, 0x20000000])
        assert step == 3                                   # three calls -> steps 1, 2, 3
        assert nbytes == 8192 + 2 * 256 * ((24 * 5 * 4 + 255) // 256)
    for it in range(3):
        X = np.concatenate([r[4][it][3] for r in res])
        Y = np.concatenate([r[4][it][4] for r in res])
        v, gX, gY = mmd_oracle.mmd2_and_grads("mix_rq_dot", X, Y, False, np.float64)
        for rank, _, _, _, out in res:
            loss, gx, gy = out[it][:3]
            assert abs(loss - v) <= 1e-6 * abs(v)
            assert np.allclose(gx, 3.0 * gX[rank * 12:(rank + 1) * 12], rtol=1e-5, atol=1e-9)
            assert np.allclose(gy, 3.0 * gY[rank * 12:(rank + 1) * 12], rtol=1e-5, atol=1e-9)


def test_numa_binding_helper_is_inert_without_topology():
    """bind_to_device_numa: cpulist parsing, and no change of the affinity mask when the topology is unknown (no GPU here)."""
    sys.path.insert(0, os.path.join(ROOT, "scaled-mmd-gan_b200"))
    from smmd.distributed import _parse_cpulist, bind_to_device_numa

    assert _parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert _parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    assert bind_to_device_numa(sysfs="/nonexistent") is None
    assert os.sched_getaffinity(0) == before
