"""Opcode histogram of every kernel in libb200rec.so (cuobjdump -sass): which kernels carry tcgen05 / TMA / TMEM code
(UTCHMMA, UTMALDG, LDTM, UTCBAR), 128-bit gathers (LDG.E.128), vector reductions (REDG), and how big they are.
usage: python profiles/sass_summary.py > profiles/r02_sass_summary.txt   (no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "inductive-recommendation_b200", "b200rec", "libb200rec.so")
KEY = ["UTCHMMA", "UTCBAR", "UTMALDG", "LDTM", "STTM", "SYNCS", "REDUX", "LDG", "LDG.E.128", "STG", "LDS", "STS", "REDG", "ATOMG", "FFMA", "FMNMX3",
       "SHFL", "BAR", "LDL", "STL", "CCTL"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
kern, rows = None, []
hist = collections.Counter()
size = 0


def flush():
    if kern is None:
        return
    name = subprocess.run(["c++filt", kern], capture_output=True, text=True).stdout.strip()
    name = re.sub(r"\(.*$", "", name).replace("b200rec::", "").replace("void ", "")
    rows.append((name, size, dict(hist)))


for line in sass.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        flush()
        kern, hist, size = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and kern:
        size = int(m.group(1), 16) + 16
        op = m.group(2)
        hist[op.split(".")[0]] += 1
        if op.startswith("LDG.E.128") or ".128" in op and op.startswith("LDG"):
            hist["LDG.E.128"] += 1
flush()
rows = [r for r in rows if not r[0].startswith("cub::") and "EmptyKernel" not in r[0]]
rows.sort(key=lambda r: -r[1])
print("# libb200rec.so -- SASS opcode counts per kernel (cuobjdump -sass, sm_100a).  size = bytes of SASS.")
print("# tcgen05.mma -> UTCHMMA, tcgen05.commit -> UTCBAR, TMA tile load -> UTMALDG, tcgen05.ld -> LDTM, tcgen05.st -> STTM, mbarrier -> SYNCS, redux.sync -> REDUX")
print("%-78s %7s " % ("kernel", "size") + " ".join("%9s" % k for k in KEY))
for name, sz, h in rows:
    print("%-78s %7d " % (name[:78], sz) + " ".join("%9d" % h.get(k, 0) for k in KEY))
print("# %d kernels (CUB sort / scan kernels of the setup-time builders not listed)" % len(rows))
