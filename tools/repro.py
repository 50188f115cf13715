import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from clair_b200 import synth, weights as W, _lib
from clair_b200.model import Clair
from oracle import clair_oracle as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
w = W.random_weights(seed=1234)
m = Clair(max_sites=max(n, 1024), batch_sites=1000)
m.set_weights(w)
X = synth.synthetic_tensors(n, seed=1000 + n)
out = m.predict_packed(X)
ref = O.forward_packed(X[:256], w, np.float64)
print("n", n, "max abs prob err (first 256)", np.abs(out[:256] - ref).max(), flush=True)
