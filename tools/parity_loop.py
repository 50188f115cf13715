"""Repeat the oracle parity check on fresh engines and several sizes (flakiness hunt): python tools/parity_loop.py [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from clair_b200 import synth, weights as W
from clair_b200.model import Clair
from oracle import clair_oracle as O
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
w = W.random_weights(seed=1234)
X = synth.synthetic_tensors(40000, seed=20240608)
ref = O.forward_packed(X[:200], w, np.float64)
ref_big_idx = np.arange(0, 40000, 211)
ref_big = O.forward_packed(X[ref_big_idx], w, np.float64)
bad = 0
for rep in range(reps):
    m = Clair(max_sites=75000, batch_sites=1000)
    m.set_weights(w)
    e1 = np.abs(m.predict_packed(X[:200]) - ref).max()
    big = m.predict_packed(X)
    e2 = np.abs(big[ref_big_idx] - ref_big).max()
    big2 = m.predict_packed(X)
    same = np.array_equal(big, big2)
    print("rep %d: 200-site err %.2e  40k-site sampled err %.2e  repeatable %s" % (rep, e1, e2, same), flush=True)
    bad += (e1 > 1e-4) + (e2 > 1e-4) + (not same)
    m.close()
print("FLAKY" if bad else "stable")
