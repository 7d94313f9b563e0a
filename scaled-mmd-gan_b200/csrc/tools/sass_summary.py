#!/usr/bin/env python
"""Per-kernel SASS opcode summary of libsmmd.so (evidence that the hot path is tcgen05 / TMEM / TMA code).
usage: sass_summary.py [lib] > profiles/r02_sass_summary.txt      (runs cuobjdump -sass; no GPU needed)"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "scaled-mmd-gan_b200/lib/libsmmd.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
kern, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kern[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        kern[cur]["_total"] += 1
        for key in ("UTCHMMA.2CTA", "UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "MUFU", "FFMA2", "FMUL2", "FADD2",
                    "SYNCS", "HMMA", "REDG", "ATOMG"):
            if op.startswith(key):
                kern[cur][key] += 1
                break
demangle = subprocess.run(["cu++filt"] + list(kern), capture_output=True, text=True).stdout.splitlines()
print("# cuobjdump -sass %s   architectures: %s" % (lib, ", ".join(arch)))
print("# SASS names: UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store,")
print("# UBLKCP = cp.async.bulk, SYNCS = mbarrier ops, FFMA2/FMUL2/FADD2 = packed fp32x2 math, HMMA = legacy mma.sync (none expected)")
cols = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "MUFU", "FFMA2", "FMUL2", "FADD2", "REDG", "HMMA", "_total"]
print("%-74s " % "kernel" + " ".join("%8s" % c.replace("UTCHMMA.2CTA", "MMA.2CTA").replace("_total", "instrs") for c in cols))
tot = collections.Counter()
for (name, c), dm in zip(kern.items(), demangle):
    short = re.sub(r"\(anonymous namespace\)::|smmd::tc::|smmd::|void ", "", dm)
    short = re.sub(r"\(.*$", "", short)
    if not any(c[k] for k in cols[:7]):
        continue   # SIMT kernels: listed in the totals only
    print("%-74s " % short[:74] + " ".join("%8d" % c[k] for k in cols))
    tot.update(c)
print("%-74s " % "TOTAL over tensor-core kernels" + " ".join("%8d" % tot[k] for k in cols))
allc = collections.Counter()
for c in kern.values():
    allc.update(c)
print("%-74s " % ("ALL %d kernels of the library" % len(kern)) + " ".join("%8d" % allc[k] for k in cols))
