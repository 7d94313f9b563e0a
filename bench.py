#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native MMD^2 / KID hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N > 1: launched by torchrun, one rank per GPU over NCCL)

Workload (BASELINE.json configs[3], the one the metric "fused MMD^2 fwd+bwd kernel-pairs/s ... % tensor
peak, target N >= 8192, d >= 256" is quoted on; configs[1] needs the conv critic, which stays in
PyTorch/cuDNN and is out of scope): mixed-RQ MMD^2 loss forward + feature gradients between N fake and N
real critic features, d = 256, synthetic data.  One "step" = one loss evaluation with both gradients.
With several GPUs the Gram is row-sharded (all_gather of the local features, fused kernel on the owned
row blocks against all columns, all_reduce of 7 partial sums): total N fixed -> "strong" scaling.

JSON line (rank 0): the driver contract (metric/value/unit/...) plus
  roofline      tensor-pipe roofline of the dominant kernel (algorithmic 14 N^2 d flop / its launch time,
                timed with CUDA events inside the library on the launching stream)
  cpu_baseline  the CPU port of the reference path on a bounded sample, on this box's host cores
  e2e           the same metric through the public Python API with HOST buffers (H2D + D2H inside the timing)
  kid           the second half of the headline metric: KID feature rows/s (configs[2])
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "scaled-mmd-gan_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

N_TOTAL = 65536        # fake rows = real rows (global): the largest N of the configs[3] sweep
D_FEAT = 256
CPU_SAMPLE_N = 4096    # cpu_baseline leg of the product arm (the reference materialises ~12 N x N fp32 temporaries per block)
KID = dict(n_codes=50000, d=2048, n_subsets=100, subset_size=1000)
KID_CPU_SUBSETS = 4
KERNEL_OF_PATH = {
    "tc_bf16_sym": "tc_sym_wgen_kernel<MathRq3Default> + tc_sym_wz_kernel (the two launches of the symmetric path, timed together)",
    "tc_bf16_symf": "tc_symf_kernel<MathRq3Default> + tc_sym_wz_kernel (fused symmetric path: Gram + W + direct products, then the mirrored products; timed together)",
    "tc_bf16_fused": "tc_fused_pair_kernel<MathRq3Default>",
    "tc_bf16_wz": "tc_wgen_kernel<MathRq3Default> + tc_wz_kernel", "tc_bf16_wz_pair": "tc_wgen_kernel<MathRq3Default, pair> + tc_wz_kernel",
}
METRIC = "mmd2_fwd_bwd_kernel_pair_evals_per_s"
UNIT = "N^2*d pair-dims/s"


def synth_features(n, d, seed, shift):
    # X = N(0,1)/sqrt(d), Y = (1.05 N(0,1) + 0.1)/sqrt(d)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=g)
    if shift:
        x = 1.05 * x + 0.1
    return (x / d ** 0.5).contiguous()


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst)"
    except Exception:
        return 1590.0, "fallback (B200_PROFILING.md 1.59 PFLOP/s)"


class ClockSampler:
    """SM clock / throttle reasons sampled through NVML (nvidia-smi's library) during the timed region."""

    def __init__(self, gpu_index, enabled=True):
        """Only rank 0 samples (its line is the one printed).  NVML is initialised HERE, before the timed region: nvmlInit / handle lookup take 100s of ms and hold
        driver locks (measured: doing it inside the region cost 6 ms/step at N = 65536); the thread only polls."""
        self.samples, self.reasons, self.query_ms = [], set(), []
        self.max_mhz = None
        self._stop = threading.Event()
        self.gpu = gpu_index
        self.nv = self.h = None
        self.enabled = enabled
        try:
            if not enabled:
                raise RuntimeError("disabled")
            import pynvml as nv

            nv.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(visible.split(",")[gpu_index]) if visible and visible.split(",")[gpu_index].isdigit() else gpu_index
            self.h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.get_reasons = (getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None)
                                or nv.nvmlDeviceGetCurrentClocksThrottleReasons)
            self.nv = nv
        except Exception as e:  # NVML missing: one nvidia-smi sample in the thread instead
            if enabled:
                self.reasons.add("nvml_unavailable:%s" % type(e).__name__)
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        if not self.enabled:
            return
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        if self.nv is not None:
            nv, h = self.nv, self.h
            while not self._stop.is_set():
                try:
                    tq = time.perf_counter()
                    self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                    r = int(self.get_reasons(h))
                    self.query_ms.append((time.perf_counter() - tq) * 1e3)
                    for nm, b in bits.items():
                        if r & b:
                            self.reasons.add(nm)
                except Exception as e:
                    self.reasons.add("nvml_error:%s" % type(e).__name__)
                    break
                self._stop.wait(0.2)    # NVML queries take driver locks: 25 Hz polling cost 14 % at 3 ms/step (8 GPUs)
            return
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=clocks.sm,clocks.max.sm",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
            a, b = [float(p) for p in out.strip().split(",")]
            self.samples.append(a)
            self.max_mhz = b
        except Exception:
            pass

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self.t.join(timeout=6)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_min_mhz": s[0] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------------------------------
# reference arm: the reference's own sources (oracle/_ref, staged by oracle/make_ref.py) on the host cores
# ----------------------------------------------------------------------------------------------------
def _reference_kind():
    """("reference", loader) when the staged reference sources (or /root/reference) are present, else ("port", None)."""
    try:
        from oracle import ref_loader
        if ref_loader.reference_available():
            return "reference", ref_loader
    except Exception:
        pass
    return "port", None


def cpu_step(n, d, kind=None, loader=None):
    """One mix_rq MMD^2 loss + both feature gradients on the host cores; returns seconds.  kind "reference": the
    unmodified gan/core/mmd.py executed over the torch-CPU `tf` shim (oracle/ref_loader.py), gradients by autograd
    through the executed reference code; kind "port": oracle/cpu_port.py (only when the sources are not staged)."""
    if kind is None:
        kind, loader = _reference_kind()
    X = synth_features(n, d, 1234, False).numpy()
    Y = synth_features(n, d, 1235, True).numpy()
    t0 = time.perf_counter()
    if kind == "reference":
        loader.reference_loss_and_grads("mix_rq", X, Y, biased=False, dtype_name="float32")
    else:
        from oracle import cpu_port
        cpu_port.mix_rq_fwd_bwd(X, Y)
    return time.perf_counter() - t0


def pick_cpu_sample(d, kind, loader, budget_s, n_calls):
    """Largest N in {2048 .. 16384} whose n_calls evaluations fit `budget_s` seconds (N^2 extrapolation of a 2048 probe)
    and whose ~40 live N x N fp32 arrays (3 Gram blocks x the reference's temporaries + autograd's saved copies) fit
    half of the free host memory."""
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 32 << 30
    cpu_step(1024, d, kind, loader)            # import / thread-pool warm-up
    t2k = cpu_step(2048, d, kind, loader)
    best = 2048
    for n in (4096, 8192, 16384):
        t_est = t2k * (n / 2048.0) ** 2
        if t_est * n_calls <= budget_s and 40.0 * n * n * 4 <= 0.5 * avail:
            best = n
    return best, t2k, avail


def reference_kid(loader, n_subsets):
    """The reference's own polynomial_mmd_averages (gan/compute_scores.py:211-229, numpy + sklearn/BLAS) on the C3
    codes with `n_subsets` subsets of 1000 (the reference scorer's default is 10, gan/utils/scorer.py:103-109)."""
    nc, d, m = KID["n_codes"], KID["d"], KID["subset_size"]
    rs = np.random.RandomState(1234)
    g = np.maximum(rs.standard_normal((nc, d)).astype(np.float32), 0)
    r = np.maximum(rs.standard_normal((nc, d)).astype(np.float32) + 0.02, 0)
    cs = loader.load_reference_compute_scores()
    np.random.seed(0)
    t0 = time.perf_counter()
    mmds = cs.polynomial_mmd_averages(g, r, n_subsets=n_subsets, subset_size=m, ret_var=False, output=None)
    t = time.perf_counter() - t0
    return {"metric": "kid_feature_rows_per_s", "value": 2.0 * n_subsets * m / t, "unit": "feature rows/s",
            "seconds": t, "kid_mean": float(np.mean(mmds)), "cores": os.cpu_count(), "kind": "reference",
            "sample": "gan/compute_scores.py polynomial_mmd_averages, %d of the %d subsets of %d on %dk vs %dk x %d fp32 codes"
                      % (n_subsets, KID["n_subsets"], m, nc // 1000, nc // 1000, d)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    kind, loader = _reference_kind()
    d = D_FEAT
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    if args.ref_n:
        n, t2k, avail = args.ref_n, None, None
    else:
        n, t2k, avail = pick_cpu_sample(d, kind, loader, budget_s=150.0, n_calls=steps + warmup)
    for _ in range(warmup):
        cpu_step(n, d, kind, loader)
    ts = [cpu_step(n, d, kind, loader) for _ in range(steps)]
    t = float(np.mean(ts))
    value = n * n * d / t
    what = ("the reference's own gan/core/mmd.py (_mix_rq_kernel + mmd2, unmodified, executed over a torch-CPU tf shim; "
            "gradients by autograd through it)") if kind == "reference" else "oracle/cpu_port.py (torch-CPU restatement)"
    sample = ("%s, fp32, fwd+bwd on %d+%d x %d: the largest N of {2048..16384} whose %d evaluations fit ~150 s and whose N x N "
              "fp32 temporaries fit host memory (the full workload is %d+%d: 17 GB per temporary); pair-dims/s is "
              "N^2-normalised, so the sample rate is the rate the reference would sustain if it could hold the workload"
              % (what, n, n, d, steps + warmup, N_TOTAL, N_TOTAL))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C4 large-batch Gram: mix_rq MMD^2 fwd+bwd, N=%d fake + %d real, d=%d" % (N_TOTAL, N_TOTAL, D_FEAT),
                   "l2": "n/a (CPU)", "sample_n": n, "same_config": False,
                   "note": "bounded sample of the workload (see cpu_baseline.sample)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind, "sample": sample,
                         "probe_seconds_at_2048": t2k, "host_mem_available_bytes": avail},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if kind == "reference" and not args.no_kid:
        try:
            line["kid"] = reference_kid(loader, n_subsets=min(10, KID["n_subsets"]))
        except Exception as e:   # sklearn / memory: the loss line stands on its own
            line["kid"] = {"unavailable": "%s: %s" % (type(e).__name__, e)}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist

    from smmd import _lib, compute_scores, mmd
    from smmd.distributed import kid_shard, sharded_mmd2

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl")
    lib = _lib.load()

    n, d = N_TOTAL, D_FEAT
    nl = n // world
    # each rank owns nl fake + nl real rows (rows [rank*nl, (rank+1)*nl) of the global batch)
    Xh = synth_features(n, d, 1234, False)[rank * nl:(rank + 1) * nl].pin_memory()
    Yh = synth_features(n, d, 1235, True)[rank * nl:(rank + 1) * nl].pin_memory()
    Xd, Yd = Xh.to(dev), Yh.to(dev)
    spec = mmd._mix_rq_kernel(Xd, Yd).spec
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)    # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    launches = [0]
    # N GPUs: how the features and the partial sums cross the GPUs -- "peer": inside the library's kernels over
    # peer-mapped memory (smmd_mmd2_fwd_bwd_peers), "nccl": all_gather + all_reduce around them
    px = None
    if world > 1:
        from smmd.distributed import PeerExchange, sharded_mmd2_raw, sharded_mmd2_raw_peers
        if args.exchange == "peer":
            # (a box whose ranks cannot map each other's memory falls back to the collective-based exchange -- on every
            # rank or on none -- and the line's `parallelism` says which one ran)
            try:
                px = PeerExchange(2 * nl, d, dev)
                okf = torch.ones(1, device=dev)
            except Exception as e:   # noqa: BLE001
                print("bench: peer-memory exchange unavailable on rank %d (%s: %s); using NCCL" % (rank, type(e).__name__, e),
                      file=sys.stderr)
                px, okf = None, torch.zeros(1, device=dev)
            dist.all_reduce(okf, op=dist.ReduceOp.MIN)
            if okf.item() == 0.0:
                px = None

    def device_step():
        """inputs resident in HBM.  1 GPU: the fused call.  N GPUs: one bf16 all_gather of the local blocks + fused
        kernel on the owned rows + all_reduce of the partial sums + combine (smmd.distributed.sharded_mmd2_raw)."""
        if world == 1:
            sc, dX, dY = mmd.fused_mmd2_raw(spec, Xd, Yd, biased=False, want_grad=True, precision="bf16")
            launches[0] += _lib.last_launch_count()
            return sc[_lib.S_MMD2], dX, dY
        if px is not None:   # gather + reduction inside the library's kernels over NVLink peer memory
            val, dX, dY, _ = sharded_mmd2_raw_peers(spec, Xd, Yd, px, biased=False, precision="bf16")
            launches[0] += _lib.last_launch_count()
            return val, dX, dY
        val, dX, dY, _ = sharded_mmd2_raw(spec, Xd, Yd, biased=False, precision="bf16")
        launches[0] += _lib.last_launch_count() + 1
        return val, dX, dY

    sampler = ClockSampler(local_rank, enabled=(rank == 0))   # NVML init happens here, outside the timed region
    # ---- warm-up ----
    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        val, dX, dY = device_step()    # (same retention pattern as the timed loop: the allocator sees identical requests)
    barrier()

    # ---- timed: K steps back to back, device time per step (L2 flushed between iterations; the flush sits
    #      outside the per-step event pair) ----
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches[0] = 0
    with sampler as clocks:
        barrier()
        t_wall0 = time.perf_counter()
        for i in range(args.steps):
            flush.zero_()
            ev[i][0].record()
            val, dX, dY = device_step()
            ev[i][1].record()
        barrier()
        t_wall = time.perf_counter() - t_wall0
    # ---- roofline: the dominant kernel alone, event pair recorded inside the library on the launching stream
    #      (separate short loop: reading the pair back synchronises, which must not sit in the timed region) ----
    lib.smmd_profile_enable(1)
    kern_ms, split_ms = [], []
    saved_launches = launches[0]
    for _ in range(5):
        flush.zero_()
        device_step()
        kern_ms.append(lib.smmd_profile_last_ms())
        split_ms.append(_lib.profile_last_split_ms())
    launches[0] = saved_launches
    lib.smmd_profile_enable(0)
    step_ms = [a.elapsed_time(b) for a, b in ev]
    if os.environ.get("SMMD_BENCH_DEBUG") and rank == 0:
        print("step_ms:", " ".join("%.2f" % v for v in step_ms), "| sampler query ms:",
              " ".join("%.1f" % v for v in getattr(sampler, "query_ms", [])), file=sys.stderr)
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    ms_per_step = total_ms.item() / args.steps
    value = float(n) * n * d / (ms_per_step * 1e-3)
    n_launch = launches[0]
    path = _lib.last_path()
    mmd2_val = float(val.item())

    # ---- roofline of the dominant kernel (this rank's share of the 14 N^2 d algorithmic flops) ----
    peak, peak_src = measured_peak()
    k_ms = float(np.mean([k for k in kern_ms if k is not None and k > 0]))
    flops_rank = 14.0 * n * n * d / world
    achieved = flops_rank / (k_ms * 1e-3) / 1e12
    traffic, kernel_name, kernels = None, KERNEL_OF_PATH.get(path, path), None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_call", {}).get(path) if world == 1 else None   # 1-GPU captures
    except Exception:
        pass
    if all(sp is not None for sp in split_ms):
        # two-kernel path: the event pair brackets both launches of one call; per-kernel times from the mid event.
        # Attribution of the algorithmic flops: 6 N^2 d Gram (W generation) + 8 N^2 d gradient GEMMs (O = W Z).
        t1 = float(np.mean([sp[0] for sp in split_ms])); t2 = float(np.mean([sp[1] for sp in split_ms]))
        if path == "tc_bf16_symf":   # direct products (4 N^2 d of the 8 N^2 d gradient flops) run inside the first kernel
            f1, f2, n1 = 10.0, 4.0, "tc_symf_kernel<MathRq3Default>"
        else:
            f1, f2, n1 = 6.0, 8.0, "tc_sym_wgen_kernel<MathRq3Default>"
        kernels = [{"name": n1, "ms": t1, "algorithmic_flops": f1 * n * n * d / world,
                    "frac": f1 * n * n * d / world / (t1 * 1e-3) / 1e12 / peak, "bound": "MUFU/FMA epilogue (3 MUFU + 24 packed fp32 ops per element pair); tensor pipe 20-40% busy"},
                   {"name": "tc_sym_wz_kernel", "ms": t2, "algorithmic_flops": f2 * n * n * d / world,
                    "frac": f2 * n * n * d / world / (t2 * 1e-3) / 1e12 / peak, "bound": "tensor / HBM reads of W"}]
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": kernel_name, "kernel_ms": k_ms, "peak_source": peak_src,
                "algorithmic_flops_per_launch": flops_rank,
                "frac_from_step_loop": 14.0 * n * n * d / world / (ms_per_step * 1e-3) / 1e12 / peak}
    try:   # informational: the same achieved rate against the SUSTAINED cuBLAS peak (this loop runs power-capped)
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            roofline["frac_of_sustained_peak"] = achieved / float(json.load(f)["bf16_tflops_sustained"])
    except Exception:
        pass
    if kernels:
        roofline["kernels"] = kernels

    # ---- e2e: public Python API, HOST buffers in, loss + gradients back on the host, every step ----
    # Two steps are kept in flight on two streams (double-buffered pinned result buffers), so one step's PCIe copies
    # overlap the other's kernels; every step still copies its inputs in and its loss + gradients out.
    nbuf = 2
    nhost = 3 if px is not None else nbuf    # pinned result buffers (peer loop: one more, its copy-out trails by a step)
    gXh = [torch.empty((nl, d), dtype=torch.float32).pin_memory() for _ in range(nhost)]
    gYh = [torch.empty((nl, d), dtype=torch.float32).pin_memory() for _ in range(nhost)]
    lossh = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(nhost)]
    # (N > 1: NCCL collectives are enqueued in program order on every rank, whichever stream they wait on)
    streams = [torch.cuda.Stream(dev) for _ in range(nbuf)]

    def e2e_step(i):
        b = i % nbuf
        st = streams[b]
        st.synchronize()            # the step that last used these result buffers (i - nbuf) is complete
        with torch.cuda.stream(st):
            flush.zero_()
            X = Xh.to(dev, non_blocking=True).requires_grad_(True)
            Y = Yh.to(dev, non_blocking=True).requires_grad_(True)
            K = mmd._mix_rq_kernel(X, Y)
            loss = sharded_mmd2(K, precision="bf16") if world > 1 else mmd.mmd2(K, precision="bf16")
            loss.backward()
            gXh[b].copy_(X.grad, non_blocking=True)
            gYh[b].copy_(Y.grad, non_blocking=True)
            lossh[b].copy_(loss.detach().reshape(1), non_blocking=True)   # device -> host read of the step's result

    # Peer exchange (N > 1): the same per-step work, but the PCIe copies are ordered behind the step's NVLink phase.
    # Measured on this pool's 8-GPU boxes: a device->host copy that runs next to the pull of the peers' rows stalls the
    # NVLink traffic for its whole duration (the gather phase grows from ~0.1 to ~1.4 ms per step, with NCCL's all_gather
    # just the same), while next to the kernels it is free.  So step i's inputs go up once step i-1's pull has completed,
    # and step i-1's results come down once step i's pull has (PeerExchange.last_pull_event): both overlap kernels only.
    # Two streams (the exchange allows two calls in flight), three pinned result buffers.
    cp_stream = torch.cuda.Stream(dev)
    done_ev = [None] * nhost
    state = {"pending": None, "prev_pull": None, "last_done": None}
    if px is not None:
        px.enable_pull_events()

    def copy_out(pending, after):
        b, gx, gy, lv, bwd_ev = pending
        if after is not None:
            cp_stream.wait_event(after)
        cp_stream.wait_event(bwd_ev)
        with torch.cuda.stream(cp_stream):
            gXh[b].copy_(gx, non_blocking=True)
            gYh[b].copy_(gy, non_blocking=True)
            lossh[b].copy_(lv, non_blocking=True)
            for t in (gx, gy, lv):
                t.record_stream(cp_stream)
            done_ev[b] = torch.cuda.Event()
            done_ev[b].record(cp_stream)
            state["last_done"] = done_ev[b]

    def e2e_step_peer(i):
        b = i % nhost
        st = streams[i % nbuf]
        if done_ev[b] is not None:
            done_ev[b].synchronize()     # the results that last used these host buffers (step i - nhost) have landed
        with torch.cuda.stream(st):
            flush.zero_()
            if state["prev_pull"] is not None:
                st.wait_event(state["prev_pull"])      # inputs go up next to the previous step's kernels
            X = Xh.to(dev, non_blocking=True).requires_grad_(True)
            Y = Yh.to(dev, non_blocking=True).requires_grad_(True)
            if state["last_done"] is not None:
                # this rank publishes (and its peers start pulling from it) only after its latest copy-out has left the
                # PCIe link: the stall is at the GPU that SERVES the NVLink reads while it copies to the host
                st.wait_event(state["last_done"])
            K = mmd._mix_rq_kernel(X, Y)
            loss = sharded_mmd2(K, precision="bf16", exchange=px)
            pull = px.last_pull_event
            loss.backward()
            bwd_ev = torch.cuda.Event()
            bwd_ev.record(st)
        if state["pending"] is not None:
            copy_out(state["pending"], pull)           # step i-1's results come down next to step i's kernels
        state["pending"] = (b, X.grad, Y.grad, loss.detach().reshape(1), bwd_ev)
        state["prev_pull"] = pull

    def e2e_drain():
        if state["pending"] is not None:
            copy_out(state["pending"], None)
            state["pending"] = None

    step_fn = e2e_step_peer if px is not None else e2e_step
    for i in range(4):
        step_fn(i)
    if px is not None:
        e2e_drain()
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_fn(i)
    if px is not None:
        e2e_drain()                 # every timed step's results are on the host before the clock stops
    barrier()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_val = float(n) * n * d / (t_e2e.item() / args.steps)
    e2e = {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": 2 * nl * d * 4 * world,
           "d2h_bytes_per_step": (2 * nl * d * 4 + 4) * world, "ms_per_step": t_e2e.item() / args.steps * 1e3,
           "api": "smmd.mmd.mmd2(smmd.mmd._mix_rq_kernel(G, images)).backward() on tensors copied from pinned host memory; "
                  "loss + both gradients copied back every step; %d step(s) in flight%s" % (
                      nbuf, "; copies ordered behind each step's NVLink pull (PeerExchange.last_pull_event)" if px is not None else "")}

    # ---- KID (configs[2]): 50k vs 50k x 2048, 100 subsets of 1000, subsets split across ranks ----
    kid = None if args.mmd_only else run_kid(dev, rank, world, args, compute_scores, lib, dist, peak)

    # ---- latency-bound shapes (configs[0] and the shipped YAML shape): microseconds per loss fwd+bwd ----
    small = run_small(dev, mmd, args) if (rank == 0 and world == 1 and not args.mmd_only) else None

    # ---- parity of the sharded path, visible to the driver (its GPU-test box has one GPU): every rank's owned-row
    #      gradients and the combined scalar against the SAME problem evaluated unsharded on rank 0 (outside all timed
    #      regions; the full feature sets are regenerated from the seeds) ----
    parity = None
    if world > 1:
        from smmd.distributed import sharded_mmd2_raw
        val_s, dXs, dYs, _ = sharded_mmd2_raw(spec, Xd, Yd, biased=False, precision="bf16")
        gx_all = [torch.empty_like(dXs) for _ in range(world)] if rank == 0 else None
        gy_all = [torch.empty_like(dYs) for _ in range(world)] if rank == 0 else None
        dist.gather(dXs, gx_all, dst=0)
        dist.gather(dYs, gy_all, dst=0)
        if rank == 0:
            Xf = synth_features(n, d, 1234, False).to(dev)
            Yf = synth_features(n, d, 1235, True).to(dev)
            sc1, dX1, dY1 = mmd.fused_mmd2_raw(spec, Xf, Yf, biased=False, want_grad=True, precision="bf16")
            path1 = _lib.last_path()
            ex = max(float((torch.cat(gx_all) - dX1).abs().max()), float((torch.cat(gy_all) - dY1).abs().max()))
            gmax = max(float(dX1.abs().max()), float(dY1.abs().max()))
            v1 = float(sc1[_lib.S_MMD2])
            parity = {"against": "the same N=%d+%d problem unsharded on rank 0 (path %s)" % (n, n, path1),
                      "mmd2_sharded": float(val_s), "mmd2_single": v1, "mmd2_rel_diff": abs(float(val_s) - v1) / abs(v1),
                      "grad_max_abs_diff": ex, "grad_max_abs": gmax, "grad_rel_to_max": ex / gmax,
                      "tolerance": "both are bf16 tensor-core evaluations (each within 4e-3 max|g| of the fp64 oracle); "
                                   "they differ by fp32 accumulation order and the r_i bookkeeping: expected <= 2e-3",
                      "ok": bool(ex <= 2e-3 * gmax and abs(float(val_s) - v1) <= 1e-4 * abs(v1))}
            del Xf, Yf, dX1, dY1
        torch.cuda.empty_cache()

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only): the reference's own code when staged ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        torch.set_num_threads(os.cpu_count() or 1)
        kind, loader = _reference_kind()
        cpu_step(CPU_SAMPLE_N, d, kind, loader)
        tc = min(cpu_step(CPU_SAMPLE_N, d, kind, loader) for _ in range(2))
        cpu = {"value": CPU_SAMPLE_N * CPU_SAMPLE_N * d / tc, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
               "sample": "%s on %d+%d x %d fp32, fwd + autograd bwd, best of 2; N^2-normalised sample -- the reference "
                         "materialises N x N temporaries and cannot hold N=%d"
                         % ("the reference's gan/core/mmd.py over the torch-CPU tf shim (oracle/_ref)" if kind == "reference"
                            else "oracle/cpu_port.py", CPU_SAMPLE_N, CPU_SAMPLE_N, d, n), "seconds_per_step_at_sample": tc}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": "C4 large-batch Gram: mix_rq MMD^2 fwd+bwd, N=%d fake + %d real, d=%d, alphas (.1,1,10), unbiased"
                                   % (n, n, d),
                       "parallelism": ("row-sharded Gram x%d (features pulled from peer memory over NVLink inside the operand "
                                       "preparation, 7 fp64 sums exchanged + combined in one kernel; no NCCL call on the data path)" % world
                                       if px is not None else
                                       "row-sharded Gram x%d (all_gather features, all_reduce 7 fp64 sums)" % world),
                       "l2": "flushed between timed iterations (256 MiB write, outside the per-step event pair)",
                       "path": path, "mmd2": mmd2_val},
            "clocks": clocks.summary(), "gpu_launches": n_launch, "roofline": roofline, "e2e": e2e, "kid": kid,
            "tflops_algorithmic": 14.0 * n * n * d / (ms_per_step * 1e-3) / 1e12,
            "wall_s_timed_region": t_wall,
            # rank 0's per-step device times (the headline `ms_per_step` is the mean of all K, max over ranks)
            "ms_per_step_min_median_max": [min(step_ms), sorted(step_ms)[len(step_ms) // 2], max(step_ms)],
            "ms_each_step": [round(v, 3) for v in step_ms],
        }
        if parity is not None:
            line["parity"] = parity
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if small is not None:
            line["small_batch_latency"] = small
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_small(dev, mmd, args):
    """C1 (64+64 x 16, mix_rbf sigmas {1,2,4,8,16}) and the shipped-YAML shape (64+64 x 1, rbf): these are
    launch-latency bound (<= 1 MFLOP), reported as microseconds per fwd+bwd, exact fp32 path, CUDA-graph replay."""
    from smmd import _lib

    out = {}
    for tag, d, mk in (("c1_mix_rbf_64x16", 16, lambda X, Y: mmd._mix_rbf_kernel(X, Y, sigmas=[1, 2, 4, 8, 16])),
                       ("yml_rbf_64x1", 1, lambda X, Y: mmd._rbf_kernel(X, Y))):
        X = torch.randn(64, d, generator=torch.Generator().manual_seed(1234)).to(dev)
        Y = (1.1 * torch.randn(64, d, generator=torch.Generator().manual_seed(1235)) + 0.1).to(dev)
        spec = mk(X, Y).spec
        for _ in range(5):
            mmd.fused_mmd2_raw(spec, X, Y, want_grad=True, precision="fp32")
        torch.cuda.synchronize()
        # plain stream launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 200
        e0.record()
        for _ in range(reps):
            sc, dX, dY = mmd.fused_mmd2_raw(spec, X, Y, want_grad=True, precision="fp32")
        e1.record()
        torch.cuda.synchronize()
        stream_us = e0.elapsed_time(e1) / reps * 1e3
        # CUDA-graph replay of 20 calls (the training loop's steady state)
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(3):
                mmd.fused_mmd2_raw(spec, X, Y, want_grad=True, precision="fp32")
            s.synchronize()
            with torch.cuda.graph(g, stream=s):
                for _ in range(20):
                    sc, dX, dY = mmd.fused_mmd2_raw(spec, X, Y, want_grad=True, precision="fp32")
        torch.cuda.synchronize()
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        graph_us = e0.elapsed_time(e1) / (20 * 20) * 1e3
        entry = {"us_per_loss_stream": stream_us, "us_per_loss_cuda_graph": graph_us, "path": _lib.last_path(),
                 "launches_per_loss": _lib.last_launch_count(), "mmd2": float(sc[0].item())}
        if not args.no_cpu:
            from oracle import mmd_oracle
            import numpy as np
            Xn, Yn = X.cpu().numpy(), Y.cpu().numpy()
            name, kw = ("mix_rbf", {"sigmas": [1, 2, 4, 8, 16]}) if d == 16 else ("rbf", {})
            t0 = time.perf_counter()
            for _ in range(20):
                v, _, _ = mmd_oracle.mmd2_and_grads(name, Xn, Yn, False, np.float32, **kw)
            entry["cpu_port_us_per_loss"] = (time.perf_counter() - t0) / 20 * 1e6
            entry["oracle_mmd2_f32"] = float(v)
        out[tag] = entry
    return out


def run_kid(dev, rank, world, args, compute_scores, lib, dist, peak):
    """configs[2]: polynomial_mmd_averages, 50k vs 50k x 2048 codes, 100 subsets of 1000, ret_var=False (the scorer's
    call, gan/utils/scorer.py:103-109).  `value`: codes and subset indices resident in HBM; `e2e`: the reference call
    signature with HOST codes (pinned), i.e. subset draw on the host (numpy global RNG, reference order) + H2D of both
    code matrices + the batched kernel + D2H of the 100 estimates, every call."""
    from smmd import _lib

    nc, d, S, m = KID["n_codes"], KID["d"], KID["n_subsets"], KID["subset_size"]
    gen = torch.Generator(device=dev).manual_seed(1234)
    # codes_g = relu(N(0,1)), codes_r = relu(N(0.02,1)); same seed on every rank = replicated codes
    g = torch.relu(torch.randn(nc, d, device=dev, generator=gen))
    r = torch.relu(torch.randn(nc, d, device=dev, generator=gen) + 0.02)
    np.random.seed(0)
    ig, ir = compute_scores.draw_subsets(nc, nc, S, m)      # reference draw order (g then r per subset)
    igd, ird = torch.from_numpy(ig).to(dev), torch.from_numpy(ir).to(dev)
    from smmd.distributed import kid_shard

    first, count = kid_shard(S, rank, world)
    nl = [0]

    def step():
        mm, _ = compute_scores.kid_subsets(g, r, igd, ird, var_at_m=nc, ret_var=False, first_subset=first, n_local=count)
        nl[0] += _lib.last_launch_count() + 1
        if world > 1:
            dist.all_reduce(mm)
        return mm

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    lib.smmd_profile_enable(1)
    reps = max(5, min(args.steps, 20))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    nl[0] = 0
    e0.record()
    for _ in range(reps):
        mm = step()
    e1.record()
    torch.cuda.synchronize()
    kms = lib.smmd_profile_last_ms()
    lib.smmd_profile_enable(0)
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    rows = 2.0 * S * m
    out = {"metric": "kid_feature_rows_per_s", "value": rows / (ms * 1e-3), "unit": "feature rows/s", "ms_per_call": ms,
           "calls_timed": reps, "gpu_launches": nl[0],
           "config": "polynomial_mmd_averages: %dk vs %dk x %d codes, %d subsets of %d, ret_var=False, codes resident in HBM "
                     "(codes 2 x 410 MB + a 1.7 GB gathered [hi|lo] workspace: larger than L2, no flush needed)"
                     % (nc // 1000, nc // 1000, d, S, m),
           "path": _lib.last_path(), "kid_mean": float(mm.mean().item()),
           "tflops_algorithmic_6m2d": 6.0 * m * m * d * S / (ms * 1e-3) / 1e12,
           "roofline": {"bound": "tensor", "unit": "TFLOP/s", "peak": peak,
                        "achieved": 6.0 * m * m * d * count / (kms * 1e-3) / 1e12 if kms and kms > 0 else None,
                        "kernel": "tc_macro_kernel<MathPoly3> 256x256 macro tiles (split-bf16: executes 3x the algorithmic flops; symmetric enumeration skips 25% of them)", "kernel_ms": kms}}
    if out["roofline"]["achieved"]:
        out["roofline"]["frac"] = out["roofline"]["achieved"] / peak
    # ---- e2e through the reference signature with host codes (rank 0, 1 GPU) ----
    if rank == 0 and world == 1:
        gh, rh = g.cpu().pin_memory(), r.cpu().pin_memory()
        def host_call():
            np.random.seed(0)
            return compute_scores.polynomial_mmd_averages(gh, rh, n_subsets=S, subset_size=m, ret_var=False, output=None)
        host_call()
        torch.cuda.synchronize()
        nrep = 3
        t0 = time.perf_counter()
        for _ in range(nrep):
            mmh = host_call()
        torch.cuda.synchronize()
        te = (time.perf_counter() - t0) / nrep
        t0 = time.perf_counter()
        np.random.seed(0)
        compute_scores.draw_subsets(nc, nc, S, m)
        t_draw = time.perf_counter() - t0
        out["e2e"] = {"value": rows / te, "unit": "feature rows/s", "ms_per_call": te * 1e3,
                      "h2d_bytes_per_step": 2 * nc * d * 4 + 2 * S * m * 4, "d2h_bytes_per_step": S * 8,
                      "host_subset_draw_ms": t_draw * 1e3,
                      "kid_mean": float(np.mean(np.asarray(mmh))),
                      "api": "smmd.compute_scores.polynomial_mmd_averages(codes_g, codes_r, n_subsets=100, subset_size=1000, "
                             "ret_var=False) with pinned HOST codes: the 200 np.random.choice(50000, 1000, replace=False) draws of "
                             "the reference, reproduced bit for bit from numpy's global RNG state by the library's host helper "
                             "(smmd_draw_subsets_mt19937: %.0f ms, running under the asynchronous H2D of both code matrices), "
                             "one batched kernel pass, D2H of the estimates" % (t_draw * 1e3)}
        del gh, rh
    if rank == 0 and world == 1 and not args.no_cpu:
        kind, loader = _reference_kind()
        gh, rh = g.cpu().numpy(), r.cpu().numpy()
        t0 = time.perf_counter()
        if kind == "reference":
            cs = loader.load_reference_compute_scores()
            np.random.seed(0)
            cs.polynomial_mmd_averages(gh, rh, n_subsets=KID_CPU_SUBSETS, subset_size=m, ret_var=False, output=None)
        else:
            from oracle import cpu_port
            cpu_port.kid_subsets(gh, rh, ig[:KID_CPU_SUBSETS], ir[:KID_CPU_SUBSETS])
        tc = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": 2.0 * KID_CPU_SUBSETS * m / tc, "unit": "feature rows/s", "cores": os.cpu_count(),
                               "kind": kind, "sample": "%d of the %d subsets (%s)" % (
                                   KID_CPU_SUBSETS, S, "the reference's gan/compute_scores.py polynomial_mmd_averages, numpy + sklearn/BLAS"
                                   if kind == "reference" else "numpy/BLAS restatement of compute_scores.py:211-335")}
    del g, r
    return out


def run_kid_workload(args):
    """`--workload kid`: the KID half of the headline metric as its own contract line."""
    import torch.distributed as dist

    from smmd import _lib, compute_scores

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl")
    lib = _lib.load()
    peak, _ = measured_peak()
    sampler = ClockSampler(local_rank, enabled=(rank == 0))
    with sampler as clocks:
        kid = run_kid(dev, rank, world, args, compute_scores, lib, dist, peak)
    if rank == 0:
        line = {"metric": kid["metric"], "value": kid["value"], "unit": kid["unit"], "n_gpus": world, "steps": kid["calls_timed"],
                "warmup": 3, "ms_per_step": kid["ms_per_call"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "bf16x3 (split-bf16 operands, fp32 accumulate)", "data": "synthetic",
                "config": {"workload": "C3 " + kid["config"], "parallelism": "subsets split across %d rank(s)" % world,
                           "l2": "working set (2.5 GB) exceeds L2", "path": kid["path"], "kid_mean": kid["kid_mean"]},
                "clocks": clocks.summary(), "gpu_launches": kid["gpu_launches"], "roofline": kid["roofline"]}
        for k in ("e2e", "cpu_baseline"):
            if k in kid:
                line[k] = kid[k]
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_sweep(args):
    """N in {4096..65536} x d in {256,512,1024}, mix_rq fwd+bwd, inputs resident in HBM.
    One JSON line per cell (rank 0): ms per evaluation (CUDA events, max over ranks, L2 flushed between iterations),
    algorithmic TFLOP/s (14 N^2 d) and its fraction of the measured bf16 peak.  Not the driver's headline line."""
    import torch.distributed as dist

    from smmd import _lib, mmd
    from smmd.distributed import sharded_mmd2_raw

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl")
    _lib.load()
    peak, _ = measured_peak()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ns = [int(v) for v in args.sweep_n.split(",")]
    ds = [int(v) for v in args.sweep_d.split(",")]
    for d in ds:
        for n in ns:
            nl = n // world
            Xd = synth_features(n, d, 1234, False)[rank * nl:(rank + 1) * nl].to(dev)
            Yd = synth_features(n, d, 1235, True)[rank * nl:(rank + 1) * nl].to(dev)
            spec = mmd._mix_rq_kernel(Xd, Yd).spec

            def step():
                if world == 1:
                    sc, dX, dY = mmd.fused_mmd2_raw(spec, Xd, Yd, biased=False, want_grad=True, precision="bf16")
                    return sc[_lib.S_MMD2]
                return sharded_mmd2_raw(spec, Xd, Yd, biased=False, precision="bf16")[0]

            steps = max(3, min(args.steps, int(2e15 / (14.0 * n * n * d / world)) + 1))
            for _ in range(3):
                flush.zero_()
                val = step()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            for i in range(steps):
                flush.zero_()
                ev[i][0].record()
                val = step()
                ev[i][1].record()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            tot = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tot, op=dist.ReduceOp.MAX)
            ms = tot.item() / steps
            if rank == 0:
                tf = 14.0 * n * n * d / (ms * 1e-3) / 1e12
                print(json.dumps({"sweep": "C4", "n": n, "d": d, "n_gpus": world, "ms": ms, "steps": steps,
                                  "pairs_per_s": float(n) * n * d / (ms * 1e-3), "tflops_algorithmic": tf,
                                  "frac_of_peak": tf / (peak * world), "path": _lib.last_path(), "mmd2": float(val.item())}),
                      flush=True)
            del Xd, Yd
            torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU-baseline legs")
    ap.add_argument("--no-kid", action="store_true", help="reference arm: skip the KID line")
    ap.add_argument("--ref-n", type=int, default=0, help="reference arm: fixed sample N (default: chosen from time / memory)")
    ap.add_argument("--mmd-only", action="store_true", help="quick A/B runs: skip the KID object and the latency-bound shapes")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: features / partial sums cross the GPUs inside the library's kernels over peer memory, or as NCCL collectives")
    ap.add_argument("--workload", default="mmd", choices=["mmd", "kid"], help="kid: the KID half of the metric as its own line")
    ap.add_argument("--sweep", action="store_true", help="C4 grid (N x d) instead of the headline line")
    ap.add_argument("--sweep-n", default="4096,8192,16384,32768,65536")
    ap.add_argument("--sweep-d", default="256,512,1024")
    args = ap.parse_args()
    if args.sweep:
        run_sweep(args)
    elif args.impl == "reference":
        run_reference(args)
    elif args.workload == "kid":
        run_kid_workload(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
