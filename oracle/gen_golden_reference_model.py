"""Golden forward outputs produced by the REFERENCE's own `clair/model.py`, run unmodified over the numpy stand-in for
TensorFlow in oracle/tf_standin/ (see its docstring for what that does and does not prove).  TEST INFRASTRUCTURE.

Run from the repo root (only where /root/reference exists):  python oracle/gen_golden_reference_model.py
    Clair()  ->  m.init()  ->  m.restore_parameters(<seed-1234 weight blob keyed by TF variable name>)  ->  m.predict(X)
exactly the calls `call_var.py` makes (clair/call_var.py:1404-1412).  The restore step also proves that the variable names
this repository uses (clair_b200/weights.py) are the ones the reference's scopes produce.  Stored in
tests/golden/reference_model_forward.npz: the inputs, the four probability arrays in float32 (the reference's float_type)
and in float64 (stand-in evaluated in double: the exact-arithmetic target for the fp64 oracle).
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"


def run_reference(weights_npz, X):
    sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_standin"))
    sys.path.insert(1, REFERENCE)
    from clair.model import Clair                      # the reference's class, unmodified
    m = Clair()
    m.init()
    m.restore_parameters(weights_npz)
    first = m.predict(X[:40])
    second = m.predict(X[40:])                         # ragged second call, like the last batch of a run
    assert m.prediction is second
    names = sorted(m.session.graph.variables)
    # the reference's own loss graph (focal loss per head, L2, task weights: clair/model.py:625-709, 783-805) through its
    # validate() (:968-1008; inference phase, so the dropouts are the identity)
    total = m.validate(X, labels_for(X))
    loss = np.array([total, m.gt21_loss, m.genotype_loss, m.indel_length_loss_1, m.indel_length_loss_2, m.l2_loss / 0.005], dtype=np.float64)
    m.close()
    return [np.concatenate([a, b]) for a, b in zip(first, second)], names, loss


def labels_for(X):
    """Truth labels [n,90] shaped like the reference's training bins (clair/utils.py: gt21 / genotype / two variant lengths,
    one-hot each; a few gt21 rows two-hot with 0.5 as a soft label would never occur, the heads are one-hot)."""
    n = len(X)
    rng = np.random.default_rng(99)
    Y = np.zeros((n, 90), np.float32)
    Y[np.arange(n), rng.integers(0, 21, n)] = 1
    Y[np.arange(n), 21 + rng.integers(0, 3, n)] = 1
    Y[np.arange(n), 24 + rng.integers(0, 33, n)] = 1
    Y[np.arange(n), 57 + rng.integers(0, 33, n)] = 1
    return Y


def main():
    if not os.path.isdir(REFERENCE):
        print("no /root/reference here: reference-model fixtures not regenerated")
        return
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        blob, xpath, out = sys.argv[2:5]
        probs, names, loss = run_reference(blob, np.load(xpath))
        np.savez(out, *probs, names=np.array(names), loss=loss)
        return
    sys.path.insert(0, ROOT)
    from clair_b200 import synth, weights as W
    w = W.random_weights(seed=1234)
    X = synth.synthetic_tensors(64, seed=20240607)
    X[3] = 0                                           # an empty pile-up
    X[4] *= 40                                         # deep coverage: saturated gates
    with tempfile.TemporaryDirectory() as tmp:
        blob = os.path.join(tmp, "weights.npz")
        W.save_blob(blob, w)
        xpath = os.path.join(tmp, "x.npy")
        np.save(xpath, X)
        res = {}
        for tag, env in (("f32", {}), ("f64", {"TF_STANDIN_FLOAT64": "1"})):
            out = os.path.join(tmp, tag + ".npz")
            subprocess.run([sys.executable, os.path.abspath(__file__), "--child", blob, xpath, out], check=True,
                           env=dict(os.environ, **env))
            with np.load(out) as z:
                res[tag] = [z["arr_%d" % k] for k in range(4)]
                res[tag + "_loss"] = z["loss"]
                names = z["names"].tolist()
    assert names == sorted(W.weight_shapes()), "variable names of the reference graph differ from clair_b200.weights"
    assert [a.dtype for a in res["f32"]] == [np.float32] * 4 and [a.shape[1] for a in res["f32"]] == [21, 3, 33, 33]
    np.savez_compressed(os.path.join(GOLD, "reference_model_forward.npz"), X=X.astype(np.int16),
                        probs_f32=np.concatenate(res["f32"], axis=1), probs_f64=np.concatenate(res["f64"], axis=1),
                        variable_names=np.array(names))
    # what the reference's validate() returned for (X, labels_for(X)): [total (lambda fed as 0), gt21, genotype, length 1,
    # length 2, L2 without lambda]
    np.savez_compressed(os.path.join(GOLD, "reference_model_loss.npz"), X=X.astype(np.int16), Y=labels_for(X),
                        loss_f32=res["f32_loss"], loss_f64=res["f64_loss"])
    print("reference_model_loss.npz:", res["f64_loss"])
    print("reference_model_forward.npz: %d sites, %d variables restored by name, f32 vs f64 max diff %.3g" % (
        len(X), len(names), np.abs(np.concatenate(res["f32"], axis=1) - np.concatenate(res["f64"], axis=1)).max()))


if __name__ == "__main__":
    main()
