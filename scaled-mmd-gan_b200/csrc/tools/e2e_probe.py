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
    nbuf = 2
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
