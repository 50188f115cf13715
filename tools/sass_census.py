#!/usr/bin/env python
"""Per-kernel SASS opcode census of the built library (run here, no GPU): which kernels carry tcgen05 / TMEM / TMA code.
usage: python tools/sass_census.py [clair_b200/lib/libclair_b200.so] > profiles/rNN_sass_census.txt
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = tensor-map TMA, UBLKCP = 1-D bulk copy, UTCBAR = tcgen05.commit,
MUFU = ex2/rcp, HMMA = legacy mma.sync (none expected), STL/LDL = register spills (B200_PROFILING.md)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "clair_b200", "lib", "libclair_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "MUFU", "HMMA", "SHFL", "STL", "LDL"]
kernels, cur = [], None
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = [m.group(1), collections.Counter(), 0]
        kernels.append(cur)
        continue
    m = re.search(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur is not None:
        cur[1][m.group(1)] += 1
        cur[2] += 1
print("library: %s  (%d bytes)" % (os.path.relpath(lib, ROOT), os.path.getsize(lib)))
print("%-64s %8s  %s" % ("kernel", "instrs", "  ".join("%7s" % w for w in WATCH)))
total = collections.Counter()
for (mangled, cnt, n), pretty in zip(kernels, names):
    short = re.sub(r"\(.*", "", pretty).replace("clairb::", "")[:64]
    print("%-64s %8d  %s" % (short, n, "  ".join("%7d" % cnt.get(w, 0) for w in WATCH)))
    total.update(cnt)
print("%-64s %8d  %s" % ("TOTAL", sum(k[2] for k in kernels), "  ".join("%7d" % total.get(w, 0) for w in WATCH)))
