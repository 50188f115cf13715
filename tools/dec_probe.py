import sys, os, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from clair_b200 import synth, weights as W
from clair_b200.model import Clair, pinned_empty
n = 75000
m = Clair(max_sites=n, batch_sites=1000); m.set_weights(W.random_weights(seed=1234))
X = pinned_empty((n, 33, 8, 4), np.float32); Xs = synth.synthetic_tensors(5000, seed=1)
for i in range(0, n, 5000): X[i:i+5000] = Xs
ref = (np.arange(n) % 4).astype(np.uint8)
def timeit(f, reps=6):
    f(); f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
print("predict_packed      %.3f ms" % timeit(lambda: m.predict_packed(X)))
print("predict             %.3f ms" % timeit(lambda: m.predict(X)))
print("predict_and_decide  %.3f ms" % timeit(lambda: m.predict_and_decide(X, ref)))
print("..._packed          %.3f ms" % timeit(lambda: m.predict_and_decide_packed(X, ref)))
m.set_profiling(True); m.predict_and_decide(X, ref); print(m.read_profile())
m.close()
