"""Row-sharded MMD^2 and subset-sharded KID across the GPUs of one node (one process per GPU,
``torch.distributed``; NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests of the host logic).

The reference never computes the loss across its towers (each tower's loss sees only its own 64+64
samples, gan/core/model.py:186-218,268-311; gradients are averaged on the CPU, :233-266).  BASELINE
config 5 asks for the *global-batch* loss, so this is new capability whose parity target is the
single-device oracle on the concatenated batch:

  1. ONE all_gather of each rank's [X_local ; Y_local] block, in bf16 when the tensor-core path will run (half
     the NVLink bytes; the library reads that block-interleaved layout directly: smmd_mmd2_fwd_bwd_gathered)
  2. each rank runs the fused kernel on ITS row block of the stacked Gram against all columns
     (smmd_problem.rank/world): complete gradients for its own rows, partial block sums
  3. all_reduce the 7 fp64 partial sums, then every rank forms the identical scalar (smmd_mmd2_combine)

No gradient collective is needed: dMMD2/dz_i for an owned row only needs the columns, which every rank
has after step 1.  KID shards the independent subsets across ranks and all_gathers S doubles.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_device_numa(device=None, sysfs="/sys"):
    """One process per GPU: pin this process to the CPUs of the NUMA node its GPU hangs off, so that pinned host buffers
    allocated afterwards (first touch) and the threads that feed the copies are local to that GPU's PCIe root.  With eight
    ranks staging their inputs / results through host memory, buffers that all sit on one socket push half of the copies
    across the inter-socket link.  Returns the node id, or None when the topology is unknown (single node, no sysfs
    entry, no CUDA device) -- then nothing is changed."""
    import os

    try:
        idx = torch.cuda.current_device() if device is None else torch.device(device).index
        props = torch.cuda.get_device_properties(idx)
        bdf = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        with open(os.path.join(sysfs, "bus/pci/devices", bdf, "numa_node")) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(os.path.join(sysfs, "devices/system/node/node%d/cpulist" % node)) as f:
            cpus = _parse_cpulist(f.read())
        cpus &= os.sched_getaffinity(0)       # never widen what the launcher allowed
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except (OSError, ValueError, AttributeError, RuntimeError, AssertionError):
        return None


def shard_rows(total, rank, world):
    """Rows [lo, hi) owned by `rank` -- the same split the C ABI uses (include/smmd.h smmd_problem)."""
    return total * rank // world, total * (rank + 1) // world


def _tc_eligible(spec, m, n, d, precision):
    """Mirror of the library's AUTO rule: will this global problem run on the bf16 tensor-core path?"""
    if precision in ("bf16", "fp16"):
        return True
    if precision in (None, "auto"):
        covered = spec.kernel_id in (_lib.K_DISTANCE, _lib.K_TANH_DISTANCE, _lib.K_RBF, _lib.K_MIX_RBF, _lib.K_MIX_RQ,
                                     _lib.K_TANH_MIX_RQ)
        return covered and (((m + n) >= 1024 and d >= 32) or d > 2048)   # include/smmd.h, SMMD_PREC_AUTO
    return False


def _default_local_compute(spec, gathered, Xl, Yl, m, n, biased, precision, rank, world):
    """smmd_mmd2_fwd_bwd_gathered on this rank's row block.  Returns (scalars[16] f64, dX_owned, dY_owned)."""
    import ctypes as C

    from .mmd import _as_ptr, _stream_ptr, _workspace

    lib = _lib.load()
    d = gathered.shape[1]
    prob = spec.problem(m, n, d, gathered.stride(0), gathered.stride(0), gathered.dtype, biased, precision, rank, world)
    dev = gathered.device
    Xo = Xl.detach().float().contiguous()
    Yo = Yl.detach().float().contiguous()
    with torch.cuda.device(dev):
        nbytes = lib.smmd_mmd2_workspace_bytes(C.byref(prob), 1)
        if nbytes == 0:
            raise _lib.SmmdError(-1, "smmd_mmd2_workspace_bytes", "problem rejected (shape/params)")
        ws = _workspace(nbytes, dev)
        scalars = torch.empty(_lib.NUM_SCALARS, dtype=torch.float64, device=dev)
        dX = torch.empty((m // world, d), dtype=torch.float32, device=dev)
        dY = torch.empty((n // world, d), dtype=torch.float32, device=dev)
        st = lib.smmd_mmd2_fwd_bwd_gathered(C.byref(prob), _as_ptr(gathered), gathered.stride(0), _as_ptr(Xo), _as_ptr(Yo),
                                            d, _as_ptr(scalars), _as_ptr(dX), _as_ptr(dY), _as_ptr(ws), nbytes,
                                            _stream_ptr(dev))
        _lib.check(st, "smmd_mmd2_fwd_bwd_gathered")
    return scalars, dX, dY


def _default_combine(spec, sums, m, n, d, biased, dtype):
    import ctypes as C

    from .mmd import _as_ptr, _stream_ptr

    lib = _lib.load()
    prob = spec.problem(m, n, d, d, d, dtype, biased, "fp32")
    out = torch.empty(1, dtype=torch.float64, device=sums.device)
    with torch.cuda.device(sums.device):
        st = lib.smmd_mmd2_combine(C.byref(prob), _as_ptr(sums), _as_ptr(out), _stream_ptr(sums.device))
    _lib.check(st, "smmd_mmd2_combine")
    return out[0]


def sharded_mmd2_raw(spec, Xl, Yl, biased=False, precision=None, group=None, local_compute=None, combine=None):
    """One sharded evaluation: ONE all_gather of this rank's [X_local ; Y_local] block (bf16 when the tensor-core
    path will run: half the NVLink bytes), the fused kernel on the owned row block, all_reduce of the partial
    sums.  Returns (mmd2 f64 scalar tensor, dX_local, dY_local, summed scalars)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    ml, nl, d = Xl.shape[0], Yl.shape[0], Xl.shape[1]
    m, n = ml * world, nl * world
    # bf16 tier: gather bf16 rows (the operand format, half the bytes).  fp16 tier: gather fp32 -- the C ABI takes f32 / bf16
    # sources only, and rounding to bf16 first would throw away the three extra bits that tier exists for.
    gdtype = torch.bfloat16 if (Xl.is_cuda and precision != "fp16" and _tc_eligible(spec, m, n, d, precision)) else torch.float32
    # this rank's block goes straight into its slot of the gather buffer (one converting copy per set, no cat + cast)
    gathered = torch.empty((world * (ml + nl), d), dtype=gdtype, device=Xl.device)
    local = gathered[rank * (ml + nl):(rank + 1) * (ml + nl)]
    local[:ml].copy_(Xl.detach())
    local[ml:].copy_(Yl.detach())
    dist.all_gather_into_tensor(gathered, local, group=group)
    sums, dX, dY = (local_compute or _default_local_compute)(spec, gathered, Xl, Yl, m, n, biased, precision, rank, world)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)   # 16 doubles; entries 1..7 are additive
    value = (combine or _default_combine)(spec, sums, m, n, d, biased, gdtype)
    return value, dX, dY, sums


class PeerExchange:
    """Exchange buffers for the peer-memory variant of the sharded loss (``smmd_mmd2_fwd_bwd_peers``): one buffer per
    rank, mapped into every other rank of the NVLink domain through torch symmetric memory (plumbing only: the data
    moves inside the library's own kernels).  Create it ONCE per (group, local shape) -- the rendezvous is a collective --
    and pass it to ``sharded_mmd2(..., exchange=px)``; every rank must then make the same sequence of calls on it.

    ``device_steps=True``: step numbers are kept on the device (CUDA-graph capturable; calls must be stream-ordered).
    ``map_buffers`` (tests): callable(nbytes, device, group) -> (own uint8 tensor, [base address of every rank]).
    """

    def __init__(self, rows_local, d, device, group=None, map_buffers=None, _compute=None, device_steps=False):
        import ctypes as C

        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > _lib.MAX_PEERS:
            raise ValueError("PeerExchange supports up to %d ranks" % _lib.MAX_PEERS)
        self.rows_local, self.d = int(rows_local), int(d)
        lib = _lib.load()
        self.nbytes = int(lib.smmd_peer_buffer_bytes(self.rows_local, self.d))
        if map_buffers is None:
            map_buffers = _symmetric_buffers
        self.buffer, ptrs, self._keepalive = map_buffers(self.nbytes, torch.device(device), group)
        self.table = _lib.PeerTable()
        self.table.world, self.table.rank = self.world, self.rank
        for r in range(self.world):
            self.table.base[r] = C.c_void_p(int(ptrs[r]))
        self.step = 0
        # device_steps: the kernels count the steps themselves (smmd_mmd2_fwd_bwd_peers with step = 0): no host state in
        # the call, so a sequence of calls can be captured into a CUDA graph; calls must then be stream-ordered
        self.device_steps = bool(device_steps)
        self._cache = {}
        self._compute = _compute     # injection point for the CPU tests of the host-side logic
        self._pull_events = None
        self.last_pull_event = None

    def enable_pull_events(self, n=4):
        """From now on every call records a CUDA event (``last_pull_event``, a rotation of `n`) on its stream once the
        peers' rows have been pulled.  A pipelined caller orders PCIe copies behind it (``stream.wait_event``) so that they
        overlap the step's kernels instead of its NVLink phase -- see smmd_peer_set_pull_event in include/smmd.h."""
        evs = []
        for _ in range(n):
            e = torch.cuda.Event()
            e.record()               # creates the underlying cudaEvent_t
            evs.append(e)
        torch.cuda.current_stream().synchronize()
        self._pull_events = evs

    def next_step(self):
        self.step += 1
        return self.step

    def fits(self, rows_local, d):
        return int(_lib.load().smmd_peer_buffer_bytes(int(rows_local), int(d))) <= self.nbytes


def _symmetric_buffers(nbytes, device, group):
    """Allocate + zero the own exchange buffer in symmetric memory and rendezvous (collective over `group`)."""
    import torch.distributed._symmetric_memory as symm_mem

    buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
    hdl = symm_mem.rendezvous(buf, group if group is not None else dist.group.WORLD)
    buf.zero_()
    torch.cuda.synchronize(device)
    hdl.barrier()            # every buffer is zero before anybody raises a flag in it
    torch.cuda.synchronize(device)
    return buf, [int(p) for p in hdl.buffer_ptrs], hdl


def _peer_local_compute(spec, px, Xl, Yl, m, n, biased, precision):
    """smmd_mmd2_fwd_bwd_peers: publish + pull + kernels + sum exchange + combine, no collective call on the data path.
    Returns (combined scalars[16] f64, dX_local, dY_local)."""
    import ctypes as C

    from .mmd import _as_ptr, _stream_ptr, _workspace

    lib = _lib.load()
    d = Xl.shape[1]
    Xo = Xl.detach()
    Yo = Yl.detach()
    if Xo.dtype != torch.float32 or not Xo.is_contiguous():
        Xo = Xo.float().contiguous()
    if Yo.dtype != torch.float32 or not Yo.is_contiguous():
        Yo = Yo.float().contiguous()
    dev = Xl.device
    key = (spec._key, m, n, d, bool(biased), precision)   # value key: kernel handles build a fresh spec per call
    ent = px._cache.get(key)
    with torch.cuda.device(dev):
        if ent is None:    # (latency-bound shapes: the struct, its byref and the workspace size are built once per signature)
            prob = spec.problem(m, n, d, d, d, torch.float32, biased, precision, px.rank, px.world)
            nbytes = lib.smmd_mmd2_workspace_bytes(C.byref(prob), 1)
            if nbytes == 0:
                raise _lib.SmmdError(-1, "smmd_mmd2_workspace_bytes", "problem rejected (shape/params)")
            ent = px._cache[key] = (prob, C.byref(prob), nbytes, C.byref(px.table))
        prob, prob_ref, nbytes, table_ref = ent
        ws = _workspace(nbytes, dev)
        scalars = torch.empty(_lib.NUM_SCALARS, dtype=torch.float64, device=dev)
        dX = torch.empty((m // px.world, d), dtype=torch.float32, device=dev)
        dY = torch.empty((n // px.world, d), dtype=torch.float32, device=dev)
        step = px.next_step()
        if px._pull_events is not None:
            px.last_pull_event = px._pull_events[step % len(px._pull_events)]
            lib.smmd_peer_set_pull_event(px.last_pull_event.cuda_event)
        st = lib.smmd_mmd2_fwd_bwd_peers(prob_ref, table_ref, 0 if px.device_steps else step, Xo.data_ptr(), Yo.data_ptr(), d,
                                         scalars.data_ptr(), dX.data_ptr(), dY.data_ptr(), ws.data_ptr(), nbytes,
                                         _stream_ptr(dev))
        if px._pull_events is not None:
            lib.smmd_peer_set_pull_event(None)
        if st != 0:
            px.step -= 1      # nothing was launched (every rank refuses the same problem): the sequence number is not used up
            _lib.check(st, "smmd_mmd2_fwd_bwd_peers")
    return scalars, dX, dY


def sharded_mmd2_raw_peers(spec, Xl, Yl, px, biased=False, precision=None):
    """One sharded evaluation over peer memory.  Returns (mmd2 f64 scalar tensor, dX_local, dY_local, combined scalars)."""
    ml, nl, d = Xl.shape[0], Yl.shape[0], Xl.shape[1]
    if not px.fits(ml + nl, d):
        raise ValueError("PeerExchange was created for %d x %d local rows; got %d x %d" % (px.rows_local, px.d, ml + nl, d))
    sums, dX, dY = (px._compute or _peer_local_compute)(spec, px, Xl, Yl, ml * px.world, nl * px.world, biased, precision)
    return sums[_lib.S_MMD2], dX, dY, sums


class _ShardedMMD2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X_local, Y_local, spec, biased, precision, group, local_compute, combine, exchange=None):
        if exchange is not None:
            value, dX, dY, sums = sharded_mmd2_raw_peers(spec, X_local, Y_local, exchange, biased, precision)
        else:
            value, dX, dY, sums = sharded_mmd2_raw(spec, X_local, Y_local, biased, precision, group, local_compute, combine)
        ctx.save_for_backward(dX, dY)
        ctx.in_dtypes = (X_local.dtype, Y_local.dtype)
        ctx.nonfinite = sums[_lib.S_NONFINITE]
        return value.to(torch.float32)

    @staticmethod
    def backward(ctx, grad_out):
        dX, dY = ctx.saved_tensors
        g = grad_out.to(dX.dtype)
        return (g * dX).to(ctx.in_dtypes[0]), (g * dY).to(ctx.in_dtypes[1]), None, None, None, None, None, None, None


def check_equal_shards(ml, nl, device, group=None):
    """All ranks must contribute equally sized blocks; raises ValueError on every rank otherwise."""
    t = torch.tensor([ml, nl, -ml, -nl], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    mx_m, mx_n, mn_m, mn_n = int(t[0]), int(t[1]), -int(t[2]), -int(t[3])
    if mx_m != mn_m or mx_n != mn_n:
        raise ValueError("sharded_mmd2: ranks hold different numbers of rows (fake %d..%d, real %d..%d)"
                         % (mn_m, mx_m, mn_n, mx_n))


def sharded_mmd2(K, biased=False, precision=None, group=None, _local_compute=None, _combine=None, check_sizes=False,
                 exchange=None):
    """Global-batch ``mmd2(kernel(G, images))`` where ``K = mmd._<name>_kernel(G_local, images_local)`` holds
    this rank's rows.  Returns the same scalar on every rank; backward yields gradients for the local rows.

    Requirements / conventions:
      * every rank passes the same number of fake rows and the same number of real rows (the gather buffer is
        `world` equal blocks; set ``check_sizes=True`` to all_reduce (min, max) the local counts once and raise on a
        mismatch instead of letting the collective hang or mis-size the problem);
      * the backward returns d(GLOBAL loss)/d(local rows) on every rank -- the gradient of the one scalar all ranks
        share, not of a per-rank average.  When the critic's parameter gradients are then averaged by DDP (all_reduce
        / world), the result is (1 / world) x the true parameter gradient: scale the loss by `world` or use a SUM
        reduction of parameter gradients to reproduce single-device training.

    ``exchange``: a :class:`PeerExchange` -- the gather of the features and the reduction of the partial sums then run
    inside the library's kernels over NVLink peer memory (no NCCL call on the data path; one launch in total for
    latency-bound shapes).  Combinations the peer entry point does not cover (mid-size exact-path problems) raise
    ``SmmdError``; call without ``exchange`` for those.

    ``_local_compute`` / ``_combine`` are injection points for the CPU tests of the host-side logic."""
    from .mmd import KernelHandle

    if not isinstance(K, KernelHandle):
        raise TypeError("sharded_mmd2 expects the handle returned by a _<name>_kernel(X_local, Y_local) call")
    if check_sizes:
        check_equal_shards(K.X.shape[0], K.Y.shape[0], K.X.device, group)
    return _ShardedMMD2.apply(K.X, K.Y, K.spec, bool(biased), precision, group, _local_compute, _combine, exchange)


def kid_shard(n_subsets, rank, world):
    """Contiguous block of subsets owned by `rank`: (first, count)."""
    lo, hi = shard_rows(n_subsets, rank, world)
    return lo, hi - lo


def sharded_polynomial_mmd_averages(codes_g, codes_r, idx_g, idx_r, ret_var=True, group=None, precision=None,
                                    _local_kid=None, **kernel_args):
    """KID over subsets split across ranks.  codes are replicated on every rank (device tensors), idx_* are
    the [S, m] subset indices drawn ONCE (rank 0's numpy RNG, reference order) and broadcast by the caller.
    Returns (mmd2[S], var[S] | None) identical on every rank."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    S = idx_g.shape[0]
    first, count = kid_shard(S, rank, world)
    m = min(codes_g.shape[0], codes_r.shape[0])
    if _local_kid is None:
        from .compute_scores import kid_subsets as _local_kid
    if count > 0:
        mm, vv = _local_kid(codes_g, codes_r, idx_g, idx_r, var_at_m=m, ret_var=ret_var, precision=precision,
                            first_subset=first, n_local=count, **kernel_args)
    else:
        mm = torch.zeros(S, dtype=torch.float64, device=idx_g.device)
        vv = torch.zeros(S, dtype=torch.float64, device=idx_g.device) if ret_var else None
    # every rank wrote only its own entries (others are zero): a SUM all-reduce assembles the vector
    dist.all_reduce(mm, op=dist.ReduceOp.SUM, group=group)
    if ret_var:
        dist.all_reduce(vv, op=dist.ReduceOp.SUM, group=group)
    return mm, vv
