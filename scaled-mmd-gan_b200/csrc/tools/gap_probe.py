import os, sys, time, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/scaled-mmd-gan_b200")
import bench
from smmd import _lib, mmd
dev = torch.device("cuda:0")
n, d = 32768, 256
X = bench.synth_features(n, d, 1234, False).to(dev); Y = bench.synth_features(n, d, 1235, True).to(dev)
spec = mmd._mix_rq_kernel(X, Y).spec
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
lib = _lib.load()
def step(): return mmd.fused_mmd2_raw(spec, X, Y, want_grad=True, precision="bf16")
for _ in range(5): step()
torch.cuda.synchronize()
for mode in ("noflush", "flush", "flush+sleep"):
    evs = []
    for i in range(10):
        if mode != "noflush": flush.zero_()
        if mode == "flush+sleep": torch.cuda._sleep(2000000)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step(); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
    print(mode, "ms/step", sum(a.elapsed_time(b) for a, b in evs) / len(evs))
# host time of one call
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): step()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("host enqueue ms/call", (t1 - t0) / 10 * 1e3, "total", (t2 - t0) / 10 * 1e3)
