import sys, os
import numpy as np
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
from clair_b200 import weights as W
from clair_b200.train import Trainer
from oracle import train_oracle as TO
from test_gpu_train import batch
w0 = W.random_weights(seed=1234)
n = 16
X1, Y1 = batch(n, 91); m1 = TO.make_masks(n, seed=1)
w1 = TO.train_step(X1, Y1, w0, m1, step=1)["new_weights"]
X, Y = batch(n, 92); masks = TO.make_masks(n, seed=2)
for label, w in (("w0", w0), ("w1", w1)):
    want = TO.train_step(X, Y, w, masks)
    t = Trainer(max_batch=16); t.set_weights({k: np.asarray(v, np.float32) for k, v in w.items()})
    t.forward_backward(X, Y, masks); t.backward_lstm()
    got = t.gradients()
    print(label)
    for k, g in want["grads"].items():
        go = g - (0.005 * np.asarray(w[k], np.float64) if "bias" not in k else 0.0)
        d = np.abs(got[k] - go)
        r = d.max() / np.abs(go).max()
        if r > 2e-5 and "L3" not in k:
            idx = np.argsort(d.ravel())[::-1][:4]
            print("  %-70s rel %.2e" % (k[-70:], r), [(tuple(int(v) for v in np.unravel_index(i, d.shape)), float(d.ravel()[i])) for i in idx])
    t.close()
