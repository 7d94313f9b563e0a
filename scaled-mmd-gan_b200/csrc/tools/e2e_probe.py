"""Multi-GPU e2e breakdown (torchrun): per phase, ms per step (max over ranks) of the host-buffer loop bench.py times.
Usage: torchrun --nproc-per-node N e2e_probe.py [--bind]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scaled-mmd-gan_b200"))
import torch, torch.distributed as dist

def main():
    bind = "--bind" in sys.argv
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
    from smmd import mmd
    from smmd.distributed import sharded_mmd2, bind_to_device_numa
    node = bind_to_device_numa(dev) if bind else None
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if world > 1: dist.init_process_group("nccl")
    import bench
    n, d = bench.N_TOTAL, bench.D_FEAT; nl = n // world
    Xh = bench.synth_features(n, d, 1234, False)[rank * nl:(rank + 1) * nl].pin_memory()
    Yh = bench.synth_features(n, d, 1235, True)[rank * nl:(rank + 1) * nl].pin_memory()
    props = torch.cuda.get_device_properties(lr)
    print("rank %d gpu %04x:%02x numa_bound=%s cpus=%d" % (rank, props.pci_domain_id, props.pci_bus_id, node, len(os.sched_getaffinity(0))), flush=True)
    def barrier():
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
    def timed(fn, steps=20, warm=4):
        for i in range(warm): fn(i)
        barrier(); t0 = time.perf_counter()
        for i in range(steps): fn(i)
        barrier()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item() / steps * 1e3
    nbuf = 3 if "--nbuf3" in sys.argv else 2
    d2hs = "--d2hstream" in sys.argv
    cp_stream = torch.cuda.Stream(dev)
    done_ev = [None] * 8
    gXh = [torch.empty((nl, d)).pin_memory() for _ in range(nbuf)]; gYh = [torch.empty((nl, d)).pin_memory() for _ in range(nbuf)]
    lossh = [torch.empty(1).pin_memory() for _ in range(nbuf)]
    streams = [torch.cuda.Stream(dev) for _ in range(nbuf)]
    Xd, Yd = Xh.to(dev), Yh.to(dev)
    def loss_of(X, Y):
        K = mmd._mix_rq_kernel(X, Y)
        return sharded_mmd2(K, precision="bf16") if world > 1 else mmd.mmd2(K, precision="bf16")
    def h2d(i):
        with torch.cuda.stream(streams[i % nbuf]):
            Xh.to(dev, non_blocking=True); Yh.to(dev, non_blocking=True)
    def d2h(i):
        b = i % nbuf
        with torch.cuda.stream(streams[b]):
            gXh[b].copy_(Xd, non_blocking=True); gYh[b].copy_(Yd, non_blocking=True)
    def compute(i):
        b = i % nbuf
        with torch.cuda.stream(streams[b]):
            X = Xd.detach().requires_grad_(True); Y = Yd.detach().requires_grad_(True)
            loss_of(X, Y).backward()
    def host_only(i):   # python/launch cost: the same calls, never waiting for the device
        compute(i)
    def full(i, sync=True):
        b = i % nbuf; st = streams[b]
        if sync: st.synchronize()
        with torch.cuda.stream(st):
            X = Xh.to(dev, non_blocking=True).requires_grad_(True); Y = Yh.to(dev, non_blocking=True).requires_grad_(True)
            loss = loss_of(X, Y); loss.backward()
            gXh[b].copy_(X.grad, non_blocking=True); gYh[b].copy_(Y.grad, non_blocking=True)
            lossh[b].copy_(loss.detach().reshape(1), non_blocking=True)
    def full_1stream(i):
        st = streams[0]
        with torch.cuda.stream(st):
            X = Xh.to(dev, non_blocking=True).requires_grad_(True); Y = Yh.to(dev, non_blocking=True).requires_grad_(True)
            loss = loss_of(X, Y); loss.backward()
            gXh[0].copy_(X.grad, non_blocking=True); gYh[0].copy_(Y.grad, non_blocking=True)
            lossh[0].copy_(loss.detach().reshape(1), non_blocking=True)
    # ---- timeline of the two-stream loop: CUDA events per phase, times relative to a base event (rank 0 prints) ----
    if "--timeline" in sys.argv:
        evs = []
        def ev():
            e = torch.cuda.Event(enable_timing=True); e.record(); return e
        marks = []
        _ag, _ar = dist.all_gather_into_tensor, dist.all_reduce
        def ag(*a, **k):
            e_a = ev(); r = _ag(*a, **k); e_b = ev(); marks.append(("ag", e_a, e_b)); return r
        def ar(*a, **k):
            e_a = ev(); r = _ar(*a, **k); e_b = ev(); marks.append(("ar", e_a, e_b)); return r
        if world > 1:
            dist.all_gather_into_tensor, dist.all_reduce = ag, ar
        def full_tl(i):
            b = i % nbuf; st = streams[b]
            if d2hs:
                if done_ev[b] is not None: done_ev[b].synchronize()
            else:
                st.synchronize()
            t_host = time.perf_counter()
            with torch.cuda.stream(st):
                e0 = ev()
                X = Xh.to(dev, non_blocking=True).requires_grad_(True); Y = Yh.to(dev, non_blocking=True).requires_grad_(True)
                e1 = ev()
                loss = loss_of(X, Y)
                e2 = ev()
                loss.backward()
                e3 = ev()
                if d2hs:
                    gx, gy, lv = X.grad, Y.grad, loss.detach().reshape(1)
                    cp_stream.wait_event(e3)
                    with torch.cuda.stream(cp_stream):
                        gXh[b].copy_(gx, non_blocking=True); gYh[b].copy_(gy, non_blocking=True); lossh[b].copy_(lv, non_blocking=True)
                        gx.record_stream(cp_stream); gy.record_stream(cp_stream); lv.record_stream(cp_stream)
                        e4 = ev()
                    done_ev[b] = e4
                else:
                    gXh[b].copy_(X.grad, non_blocking=True); gYh[b].copy_(Y.grad, non_blocking=True)
                    lossh[b].copy_(loss.detach().reshape(1), non_blocking=True)
                    e4 = ev()
            evs.append((i, t_host, time.perf_counter(), e0, e1, e2, e3, e4))
        for i in range(6): full_tl(i)
        evs.clear(); marks.clear()
        barrier()
        base = torch.cuda.Event(enable_timing=True); base.record(); tb = time.perf_counter()
        t0 = time.perf_counter()
        for i in range(12): full_tl(i)
        barrier()
        dist.all_gather_into_tensor, dist.all_reduce = _ag, _ar
        if rank == 0:
            print("TL variant nbuf=%d d2hstream=%s: %.3f ms/step" % (nbuf, d2hs, (time.perf_counter() - t0) / 12 * 1e3), flush=True)
            for (nm, e_a, e_b) in marks[:24]:
                print("TL   %s  [%.2f .. %.2f]" % (nm, base.elapsed_time(e_a), base.elapsed_time(e_b)), flush=True)
            for (i, th0, th1, e0, e1, e2, e3, e4) in evs:
                print("TL step %d host[%.2f..%.2f] start %.2f | h2d_end %.2f | fwd_end %.2f | bwd_end %.2f | d2h_end %.2f" % (
                    i, (th0 - tb) * 1e3, (th1 - tb) * 1e3, base.elapsed_time(e0), base.elapsed_time(e1), base.elapsed_time(e2),
                    base.elapsed_time(e3), base.elapsed_time(e4)), flush=True)
    if "--timeline" in sys.argv:
        if world > 1: dist.destroy_process_group()
        return
    res = {}
    for name, fn in (("h2d", h2d), ("d2h", d2h), ("compute", compute), ("full_2streams", full), ("full_1stream", full_1stream)):
        res[name] = timed(fn)
    # host enqueue time of one compute step (no device wait inside)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(10): compute(i)
    res["host_enqueue_compute"] = (time.perf_counter() - t0) / 10 * 1e3
    barrier()
    if rank == 0: print("E2E_PROBE world=%d bind=%s " % (world, bind) + " ".join("%s=%.3f" % kv for kv in res.items()), flush=True)
    if world > 1: dist.destroy_process_group()
main()
