import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
from clair_b200 import synth, weights as W
from clair_b200.train import Trainer
from oracle import train_oracle as TO
from test_gpu_train import batch
w0 = W.random_weights(seed=1234)
n = 16
t = Trainer(max_batch=16); t.set_weights(w0)
w, state = w0, None
K = 'LSTM1/stack_bidirectional_rnn/cell_0/bidirectional_rnn/fw/cudnn_compatible_lstm_cell/kernel'
for step in (1, 2):
    X, Y = batch(n, 90 + step)
    masks = TO.make_masks(n, seed=step)
    want = TO.train_step(X, Y, w, masks, adam_state=state, step=step)
    t.forward_backward(X, Y, masks); t.backward_lstm()
    gd = t.gradients()[K].astype(np.float64)
    go = want["grads"][K] - 0.005 * np.asarray(w[K], np.float64)
    print("step", step, "grad rel err (max abs / max abs)", np.abs(gd - go).max() / np.abs(go).max(), "max|g|", np.abs(go).max())
    t.grad_norm = t.apply()
    got = t.get_weights()[K]
    d = np.abs(got - want["new_weights"][K])
    g = want["grads"][K]
    solid = np.abs(g) > 1e-3 * np.abs(g).max()
    dd = np.where(solid, d, 0)
    i = np.unravel_index(dd.argmax(), dd.shape)
    print(" worst solid", i, dd[i], "g(oracle, with L2)", g[i], "dev grad(no L2)", gd[i], "oracle grad(no L2)", go[i], "w", w[K][i])
    if state is not None:
        print(" oracle m,v before", state[K][0][i], state[K][1][i])
    print(" dev m,v after", t._get(2)[K][i], t._get(3)[K][i], "oracle m,v after", want["adam_state"][K][0][i], want["adam_state"][K][1][i], "norm dev/oracle", t.grad_norm, want["grad_norm"])
    w, state = want["new_weights"], want["adam_state"]
    t.set_weights({k: np.asarray(v, np.float32) for k, v in w.items()})
