"""Print the l3l4_fused timeline probe (CLAIRB_LF_TRACE=<file> python tools/kbench.py ...): per channel c of tile 0,
cycles relative to channel c0's producer stamp.  P = producer got the stage, L3 = stage landed and D3 free: L3 MMAs issued, E8 = epilogue saw D3, E9 = epilogue math done, E10 = A4 buffer free, M4 = L4 MMAs issued."""
import sys
import numpy as np
t = np.loadtxt(sys.argv[1], dtype=np.int64)
c0 = int(sys.argv[2]) if len(sys.argv) > 2 else 100
n = int(sys.argv[3]) if len(sys.argv) > 3 else 8
base = t[c0][0]
names = {0: "P", 2: "L3", 3: "L3i", 4: "M4", 5: "M4i", 8: "E8", 9: "E9", 10: "E10", 11: "A0", 12: "A1", 13: "A2", 14: "A3"}
for c in range(c0, c0 + n):
    print("c%3d " % c + "  ".join("%s@%d" % (names[e], t[c][e] - base) for e in sorted(names) if t[c][e]))
print("cycles per channel over 100..200: %.0f" % ((t[200][2] - t[100][2]) / 100.0))
