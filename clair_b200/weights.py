"""Weight naming, random initialisation and the flat weight-blob format.

Variable names and shapes follow the TF-1.13 graph the reference builds
(clair/model.py:265-312 LSTM cells, :225-244 slice dense, :482-618 dense layers); see
SURVEY.md section 8a for the table.  The blob is a plain ``.npz`` keyed by TF variable name.
"""
import numpy as np

T, F, H = 33, 32, 128
L3_UNITS, L4_UNITS, L5_UNITS = 30, 192, 96
HEADS = (21, 3, 33, 33)                          # clair/task/main.py:10-29
HEAD_NAMES = ("Y_base_change_logits", "Y_genotype_logits",
              "Y_indel_length_logits_1", "Y_indel_length_logits_2")   # clair/model.py:581-618
N_OUT = sum(HEADS)

_LSTM = "{layer}/stack_bidirectional_rnn/cell_0/bidirectional_rnn/{d}/cudnn_compatible_lstm_cell/{v}"


def lstm_name(layer, direction, var):
    return _LSTM.format(layer="LSTM%d" % layer, d=direction, v=var)


def weight_shapes():
    """Ordered {tf_name: shape} of every trainable variable on the forward path."""
    s = {}
    for layer, fin in ((1, F), (2, 2 * H)):
        for d in ("fw", "bw"):
            s[lstm_name(layer, d, "kernel")] = (fin + H, 4 * H)
            s[lstm_name(layer, d, "bias")] = (4 * H,)
    for c in range(2 * H):
        s["L3/Unit_%d/kernel" % c] = (T, L3_UNITS)
        s["L3/Unit_%d/bias" % c] = (L3_UNITS,)
    s["L4/kernel"] = (L3_UNITS * 2 * H, L4_UNITS)
    s["L4/bias"] = (L4_UNITS,)
    for k in range(4):
        s["L5_%d/kernel" % (k + 1)] = (L4_UNITS, L5_UNITS)
        s["L5_%d/bias" % (k + 1)] = (L5_UNITS,)
    for k in range(4):
        s["Prediction/%s/kernel" % HEAD_NAMES[k]] = (L5_UNITS, HEADS[k])
        s["Prediction/%s/bias" % HEAD_NAMES[k]] = (HEADS[k],)
    return s


def n_params():
    return int(sum(np.prod(v) for v in weight_shapes().values()))


def _trunc_normal(rng, shape, std):
    out = rng.standard_normal(shape)
    bad = np.abs(out) > 2.0
    while bad.any():
        out[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(out) > 2.0
    return out * std


def random_weights(seed=1234, bias_std=0.05):
    """Random "ONT-shape" weights with the reference's initialisers.

    Dense kernels: variance_scaling_initializer(factor=1, FAN_IN) (clair/model.py:394-398)
    = truncated normal, std sqrt(1.3/fan_in).  LSTM kernels: TF default glorot_uniform.
    TF initialises biases to zero; tests use small non-zero biases so the bias path is exercised.
    """
    rng = np.random.default_rng(seed)
    w = {}
    for name, shape in weight_shapes().items():
        if name.endswith("bias"):
            w[name] = (rng.standard_normal(shape) * bias_std).astype(np.float32)
        elif "lstm_cell" in name:
            limit = np.sqrt(6.0 / (shape[0] + shape[1]))
            w[name] = rng.uniform(-limit, limit, shape).astype(np.float32)
        else:
            w[name] = _trunc_normal(rng, shape, np.sqrt(1.3 / shape[0])).astype(np.float32)
    return w


def save_blob(path, weights):
    np.savez(path, **{k: np.asarray(v, dtype=np.float32) for k, v in weights.items()})


def load_blob(path):
    if not str(path).endswith(".npz"):
        path = str(path) + ".npz"
    with np.load(path) as z:
        w = {k: np.asarray(z[k], dtype=np.float32) for k in z.files}
    check_weights(w)
    return w


def check_weights(w):
    shapes = weight_shapes()
    missing = [k for k in shapes if k not in w]
    if missing:
        raise ValueError("weight blob is missing %d variables, e.g. %s" % (len(missing), missing[0]))
    for k, s in shapes.items():
        if tuple(w[k].shape) != tuple(s):
            raise ValueError("weight %s has shape %s, expected %s" % (k, tuple(w[k].shape), s))
