"""One small training step (and one create_tensors call) for compute-sanitizer runs:
compute-sanitizer --tool memcheck|racecheck python tools/probes/train_once.py"""
import os, sys
import numpy as np
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
from clair_b200 import synth, weights as W
from clair_b200.train import Trainer
n = int(sys.argv[1]) if len(sys.argv) > 1 else 72
t = Trainer(max_batch=n)
t.set_weights(W.random_weights(seed=1234))
X = synth.synthetic_tensors(n, seed=3).astype(np.int16)
Y = np.zeros((n, 90), np.float32); Y[:, [0, 21, 24, 57]] = 1
print("loss", t.train(X, Y), "norm", t.grad_norm)
print("loss", t.train(X, Y), "norm", t.grad_norm)
t.close()
