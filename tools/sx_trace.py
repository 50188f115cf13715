"""Print the lstm_seq_x2 timeline probe (CLAIRB_SX_TRACE=<file> python tools/kbench.py ...): cycles relative to the
step's first event.  Issuer events I.*; epilogue (warp 1) events per block b: full, drained, math/post per slice, done."""
import sys
import numpy as np
names = {0: "I.x01.start", 1: "I.x01.issued", 2: "I.h.wait0", 3: "I.h.wait3", 4: "I.h01.committed",
         8: "I.h2.start", 9: "I.h3.start", 12: "I.x23.committed"}
if "--layer1" in sys.argv:       # CLAIRB_S1_TRACE (lstm_seq<FUSE_X>): I.acc0 .. I.end are the issuer's events
    sys.argv.remove("--layer1")
    names = {0: "I.acc0", 1: "I.acc1", 2: "I.h.wait0", 3: "I.h.wait3", 4: "I.b01.committed", 8: "I.acc2", 9: "I.acc3", 12: "I.end"}
t = np.loadtxt(sys.argv[1], dtype=np.int64)
steps = [int(x) for x in sys.argv[2:]] or [10, 11, 12]
for b in range(4):
    for k, nm in enumerate(["full", "drained", "s0", "s1", "", "", "", "done" if b == 3 else ""]):
        if nm:
            names[16 + b * 8 + k] = "E%d.%s" % (b, nm)
for s in steps:
    base = t[s][0]
    ev = sorted((t[s][e] - base, names[e]) for e in names if t[s][e] != 0)
    print("step %d (period %d):" % (s, t[s + 1][0] - t[s][0] if s + 1 < len(t) else -1))
    print("   " + "  ".join("%s@%d" % (n, c) for c, n in ev))
