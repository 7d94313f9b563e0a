"""Developer probe: the latency-bound C1 shape (64+64 x 16 mix_rbf) through the exact path, a few calls
(run under `ncu --metrics gpu__time_duration.sum` for per-kernel durations)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(ROOT, "scaled-mmd-gan_b200"))
from smmd import _lib, mmd  # noqa: E402

d = int(sys.argv[1]) if len(sys.argv) > 1 else 16
X = torch.randn(64, d, device="cuda")
Y = 1.1 * torch.randn(64, d, device="cuda") + 0.1
spec = mmd._mix_rbf_kernel(X, Y, sigmas=[1, 2, 4, 8, 16]).spec
for _ in range(4):
    sc, dX, dY = mmd.fused_mmd2_raw(spec, X, Y, want_grad=True, precision="fp32")
torch.cuda.synchronize()
print(_lib.last_path(), _lib.last_launch_count(), float(sc[0]))
