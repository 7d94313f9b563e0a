"""TEST / BENCH INFRASTRUCTURE ONLY -- stages the reference's own hot-path sources for the `--impl reference` arm.

The reference (playHing/Scaled-MMD-GAN) is pure TF-1.x Python with no build system, so "building" it means making
the three files of the path available where `oracle/ref_loader.py` can execute them on the GPU box's host cores
(`/root/reference` exists only in the build container):

    gan/core/mmd.py            whole file (kernels :18-188, mmd2 :194-220, ratio :223-293, 3-sample :296-539)
    gan/compute_scores.py      whole file (KID: :211-335; the TF feature extractors in it are never called)
    gan/core/ops.py:209-225    `sq_sum` and `dot` only (the rest of ops.py imports matplotlib / scipy.misc)

Files are copied BYTE FOR BYTE into the git-ignored `oracle/_ref/` (they travel to the GPU box with the snapshot, they
never enter the history).  Called by `__graft_entry__.build()` whenever /root/reference is present.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("SMMD_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
OPS_LINES = (209, 225)   # sq_sum + dot


def stage(src=SRC, dst=DST) -> bool:
    if not os.path.isfile(os.path.join(src, "gan", "core", "mmd.py")):
        return False
    os.makedirs(os.path.join(dst, "gan", "core"), exist_ok=True)
    shutil.copyfile(os.path.join(src, "gan", "core", "mmd.py"), os.path.join(dst, "gan", "core", "mmd.py"))
    shutil.copyfile(os.path.join(src, "gan", "compute_scores.py"), os.path.join(dst, "gan", "compute_scores.py"))
    with open(os.path.join(src, "gan", "core", "ops.py")) as f:
        lines = f.readlines()
    with open(os.path.join(dst, "gan", "core", "ops_sq_sum_dot.py"), "w") as f:
        f.writelines(lines[OPS_LINES[0] - 1:OPS_LINES[1]])
    with open(os.path.join(dst, "SOURCE.txt"), "w") as f:
        f.write("verbatim copies from %s (gan/core/mmd.py, gan/compute_scores.py, gan/core/ops.py:%d-%d)\n"
                % (src, OPS_LINES[0], OPS_LINES[1]))
    return True


if __name__ == "__main__":
    ok = stage()
    print("staged reference sources into %s" % DST if ok else "reference not found at %s" % SRC)
    sys.exit(0 if ok else 1)
