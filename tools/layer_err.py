"""Per-layer max abs error of the device path against the fp64 oracle (256 sites): python tools/layer_err.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from clair_b200 import synth, weights as W, _lib
from clair_b200.model import Clair
from oracle import clair_oracle as O
w = W.random_weights(seed=1234)
X = synth.synthetic_tensors(256, seed=20240607)
ref_probs, im = O.forward(X, w, np.float64, intermediates=True)
m = Clair(max_sites=1024, batch_sites=1000)
m.set_weights(w)
m.predict(X)
out = []
for layer, key in ((_lib.LAYER_LSTM1, "lstm1"), (_lib.LAYER_LSTM2, "lstm2"), (_lib.LAYER_L3, "l3"), (_lib.LAYER_L4, "l4")):
    got = m.get_layer(layer, 256)
    out.append("%s %.2e" % (key, np.abs(got - im[key]).max()))
lg = m.get_layer(_lib.LAYER_LOGITS, 256)
rl = np.concatenate(im["logits"], axis=1)
out.append("logits(scaled) %.2e" % (np.abs(lg - rl) / np.maximum(1, np.abs(rl))).max())
print(os.environ.get("CLAIRB_L2_STREAM", "1"), os.environ.get("CLAIRB_ENGINE", "tc"), "  ".join(out))
m.close()
