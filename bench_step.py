#!/usr/bin/env python
"""bench_step.py -- BASELINE.json configs[1] (C2) and configs[4] (C5): where the loss sits inside a training step.

  python bench_step.py [--steps K] [--warmup W]                      C2: cifar10_smmd.yml critic step on 1 B200
  torchrun --nproc-per-node 8 ... bench_step.py --config c5          C5: global-batch loss across the ranks

The conv nets are host-side PyTorch/cuDNN and out of scope; this harness exists to MEASURE the loss
inside a step.  C2: SNGAN-32 critic + generator (layer shapes of gan/core/architecture.py:211-230, 395-407, spectral
norm on the critic), batch 64, kernel rbf, dof_dim 1 exactly as configs/cifar10_smmd.yml (plus a mix_rbf / dof_dim 16
variant), SMMD scaling variant 'grad', scaling_coeff 10 (gan/main.py:101-102).  One critic step = G forward (no grad),
critic on fake + real, Jacobian-norm scale (double backward), loss, backward, Adam.  The same step is timed with
  fused : smmd.scaling.scaled_mmd2 over libsmmd's one-launch loss kernel
  eager : the reference's dense formulation as plain torch ops on the GPU (3 Grams, exp, sums; autograd backward)
and the loss alone (forward + backward on detached features) is timed both ways, stream launches and CUDA graph.
C5: every rank holds 64 fake + 64 real critic outputs (dof_dim 1, imagenet_smmd.yml); the global 512 + 512 loss is
smmd.distributed.sharded_mmd2 (all_gather + fused kernel on the owned rows + all_reduce): microseconds per loss, plain
launches and CUDA-graph replay, with the two collectives timed alone beside it.
One JSON line per config (rank 0).  Not the driver's headline line (bench.py)."""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "scaled-mmd-gan_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

SN = nn.utils.parametrizations.spectral_norm


class Critic(nn.Module):      # SNGANDiscriminator, architecture.py:395-407 (7 SN convs + SN linear), df_dim 64
    def __init__(self, dof_dim=1, size=32):
        super().__init__()
        cfg = [(3, 64, 3, 1), (64, 128, 4, 2), (128, 128, 3, 1), (128, 256, 4, 2), (256, 256, 3, 1), (256, 512, 4, 2), (512, 512, 3, 1)]
        self.convs = nn.ModuleList([SN(nn.Conv2d(i, o, k, s, 1)) for i, o, k, s in cfg])
        self.lin = SN(nn.Linear(512 * (size // 8) ** 2, dof_dim))

    def forward(self, x):
        for c in self.convs:
            x = nn.functional.leaky_relu(c(x), 0.2)
        return self.lin(x.flatten(1))


class Generator(nn.Module):   # SNGANGenerator, architecture.py:211-230, gf_dim 64, z_dim 128
    def __init__(self, size=32):
        super().__init__()
        s8 = size // 8
        self.s8 = s8
        self.lin = nn.Linear(128, 512 * s8 * s8)
        self.bn0 = nn.BatchNorm2d(512)
        self.up = nn.ModuleList([nn.ConvTranspose2d(512, 256, 4, 2, 1), nn.ConvTranspose2d(256, 128, 4, 2, 1), nn.ConvTranspose2d(128, 64, 4, 2, 1)])
        self.bns = nn.ModuleList([nn.BatchNorm2d(256), nn.BatchNorm2d(128), nn.BatchNorm2d(64)])
        self.out = nn.ConvTranspose2d(64, 3, 3, 1, 1)

    def forward(self, z):
        h = torch.relu(self.bn0(self.lin(z).view(-1, 512, self.s8, self.s8)))
        for u, b in zip(self.up, self.bns):
            h = torch.relu(b(u(h)))
        return torch.sigmoid(self.out(h))


def eager_mmd2(name, X, Y):
    """The reference's dense formulation (mmd.py:55-116, 194-220) as torch ops; X = fake, Y = real."""
    sig = [1.0] if name == "rbf" else [1.0, 2.0, 4.0, 8.0, 16.0]
    XX, XY, YY = X @ X.T, X @ Y.T, Y @ Y.T
    nx, ny = torch.diagonal(XX), torch.diagonal(YY)
    def k(G, a, b):
        D = torch.clamp(-2 * G + a[:, None] + b[None, :], min=0)
        return sum(torch.exp(-D / (2 * s * s)) for s in sig)
    m, n, cd = float(X.shape[0]), float(Y.shape[0]), float(len(sig))
    return ((k(XX, nx, nx).sum() - m * cd) / (m * (m - 1)) + (k(YY, ny, ny).sum() - n * cd) / (n * (n - 1))
            - 2 * k(XY, nx, ny).sum() / (m * n))


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3   # microseconds


def graphed(fn, reps=20):
    """microseconds per call under CUDA-graph replay of `reps` calls, or None if capture is not possible."""
    try:
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(3):
                fn()
            s.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                for _ in range(reps):
                    fn()
        torch.cuda.synchronize()
        return timed(g.replay, 20, 3) / reps
    except Exception as e:   # noqa: BLE001
        return "capture failed: %s" % type(e).__name__


def run_c2(args):
    from smmd import _lib, mmd, scaling

    dev = torch.device("cuda", 0)
    torch.manual_seed(1)
    out = {"config": "C2 cifar10_smmd.yml: SNGAN-32, batch 64, SMMD critic step (scaling 'grad', coeff 10), synthetic images U[0,1]",
           "steps": args.steps, "warmup": args.warmup}
    images = torch.rand(64, 3, 32, 32, device=dev)
    for tag, kname, dof in (("yml_rbf_dof1", "rbf", 1), ("mix_rbf_dof16", "mix_rbf", 16)):
        D, G = Critic(dof).to(dev), Generator().to(dev)
        opt = torch.optim.Adam(D.parameters(), lr=1e-4, betas=(0.5, 0.9))
        kern = (lambda a, b: mmd._rbf_kernel(a, b)) if kname == "rbf" else (lambda a, b: mmd._mix_rbf_kernel(a, b, sigmas=[1, 2, 4, 8, 16]))

        def step(fused):
            z = torch.rand(64, 128, device=dev) * 2 - 1
            with torch.no_grad():
                fake = G(z)
            x = images.detach().requires_grad_(True)
            d_img, d_fake = D(x), D(fake)
            scale, _, _ = scaling.smmd_scale(d_img, x, 10.0, "grad")
            if fused:
                g_loss, _ = scaling.scaled_mmd2(kern(d_fake, d_img), scale, precision="fp32")
            else:
                g_loss = eager_mmd2(kname, d_fake, d_img) * scale
            opt.zero_grad(set_to_none=True)
            (-g_loss).backward()
            opt.step()
            return g_loss

        f = torch.randn(64, dof, device=dev)
        r = 1.1 * torch.randn(64, dof, device=dev) + 0.1

        def loss_only(fused):
            a, b = f.detach().requires_grad_(True), r.detach().requires_grad_(True)
            v = mmd.mmd2(kern(a, b), precision="fp32") if fused else eager_mmd2(kname, a, b)
            v.backward()
            return v

        spec = kern(f, r).spec
        e = {"step_us_fused": timed(lambda: step(True), args.steps, args.warmup),
             "step_us_eager_loss": timed(lambda: step(False), args.steps, args.warmup),
             "loss_fwd_bwd_us_fused_autograd": timed(lambda: loss_only(True), 200, 20),
             "loss_fwd_bwd_us_eager": timed(lambda: loss_only(False), 200, 20),
             "loss_fwd_bwd_us_fused_raw_call": timed(lambda: mmd.fused_mmd2_raw(spec, f, r, want_grad=True, precision="fp32"), 200, 20),
             "loss_fwd_bwd_us_fused_cuda_graph": graphed(lambda: mmd.fused_mmd2_raw(spec, f, r, want_grad=True, precision="fp32")),
             "loss_path": None, "loss_launches": None}
        mmd.fused_mmd2_raw(spec, f, r, want_grad=True, precision="fp32")
        e["loss_path"], e["loss_launches"] = _lib.last_path(), _lib.last_launch_count()
        v1, v2 = float(step(True)), float(step(False))
        e["g_loss_fused_vs_eager_same_weights_next_steps"] = [v1, v2]
        e["loss_share_of_step_fused"] = e["loss_fwd_bwd_us_fused_autograd"] / e["step_us_fused"]
        e["loss_share_of_step_eager"] = e["loss_fwd_bwd_us_eager"] / e["step_us_eager_loss"]
        out[tag] = e
    print(json.dumps(out))


def run_c5(args):
    import torch.distributed as dist

    from smmd import _lib, mmd
    from smmd.distributed import PeerExchange, sharded_mmd2_raw, sharded_mmd2_raw_peers

    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl")
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    f = torch.randn(64, 1, device=dev, generator=g)
    r = 1.1 * torch.randn(64, 1, device=dev, generator=g) + 0.1
    spec = mmd._rbf_kernel(f, r).spec
    buf = torch.empty(world * 128, 1, device=dev)
    sc = torch.zeros(16, dtype=torch.float64, device=dev)

    def loss():
        return sharded_mmd2_raw(spec, f, r, biased=False, precision="fp32")[0]

    px = PeerExchange(128, 1, dev)

    def loss_peers():   # the same loss with the gather + reduction inside ONE kernel over NVLink peer memory
        return sharded_mmd2_raw_peers(spec, f, r, px, biased=False, precision="fp32")[0]

    pxg = PeerExchange(128, 1, dev, device_steps=True)   # steps counted on the device: capturable

    def loss_peers_graphable():
        return sharded_mmd2_raw_peers(spec, f, r, pxg, biased=False, precision="fp32")[0]

    def collectives_only():
        dist.all_gather_into_tensor(buf, buf[rank * 128:(rank + 1) * 128])
        dist.all_reduce(sc)

    def local_only():
        return mmd.fused_mmd2_raw(spec, f, r, want_grad=True, precision="fp32")

    res = {"config": "C5 imagenet_smmd.yml loss: %d ranks x (64 fake + 64 real) x dof_dim 1, rbf, global-batch MMD^2 "
                     "(all_gather + owned-row kernel + all_reduce of 7 sums)" % world, "n_gpus": world,
           "sharded_loss_peer_memory_us": timed(loss_peers, 200, 20), "mmd2_peer_memory": float(loss_peers()),
           "path_peer_memory": _lib.last_path(),
           "sharded_loss_peer_memory_us_cuda_graph": graphed(loss_peers_graphable), "mmd2_peer_memory_graphable": float(loss_peers_graphable()),
           "sharded_loss_us": timed(loss, 200, 20), "collectives_only_us": timed(collectives_only, 200, 20),
           "local_64x1_loss_us": timed(local_only, 200, 20), "sharded_loss_us_cuda_graph": graphed(loss),
           "collectives_only_us_cuda_graph": graphed(collectives_only), "mmd2": float(loss()), "path": _lib.last_path()}
    t = torch.tensor([res["sharded_loss_us"]], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res["sharded_loss_us_max_over_ranks"] = float(t)
    if rank == 0:
        print(json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2", choices=["c2", "c5"])
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    a = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("bench_step.py needs a B200: there is no CPU fallback for the loss")
    (run_c2 if a.config == "c2" else run_c5)(a)
