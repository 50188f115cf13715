"""CPU restatement of Clair's TRAINING step with torch autograd.  TEST INFRASTRUCTURE ONLY.

Follows the reference's graph in training mode and its loss / optimiser (clair/model.py):
    forward                  :400-622   (as oracle/clair_oracle.py) with phase_placeholder = True:
        tf.layers.dropout after LSTM1 (rate 0) and LSTM2 (rate 0.5)      :434-440, :453-459   x * mask / keep_prob
        selu.dropout_selu after L4 (0.5) and L5_1..4 (0.2)               :495-578, clair/selu.py:43-74
            ret = a * (x * mask + alpha' * (1 - mask)) + b,  alpha' = -1.7580993408473766,
            a = sqrt(1 / (q * ((1 - q) * alpha'^2 + 1))),  b = -a * (1 - q) * alpha',  q = keep_prob
    focal loss per head      :783-805   p = softmax(z);  -(where(t>0, t-p, 0)^2 log clip(p) + where(t>0, 0, p)^2 log clip(1-p)), summed
    L2                       :689-694   lambda * sum over non-bias variables of ||v||^2 / 2
    total                    :696-709   task_loss_weights . [gt21, genotype, length 1, length 2, L2]   (sums over the batch, no mean)
    clip, Adam               :717-728   clip_by_global_norm(5.0); AdamOptimizer defaults beta1 0.9, beta2 0.999, eps 1e-8
The dropout masks are inputs (TensorFlow's random stream cannot be reproduced; the device generates its own from a seed and
the parity tests hand both sides the same masks).

PARITY: the loss definition is pinned on the reference's own loss graph - tests/golden/reference_model_loss.npz holds what
the reference's unmodified `Clair.validate` returns over oracle/tf_standin (phase False: dropouts are the identity there).
Gradients and the optimiser step rest on torch autograd / the published Adam update; TensorFlow cannot run here.
"""
import numpy as np
import torch

from . import clair_oracle as O

ALPHA_DROPOUT = -1.7580993408473766                   # clair/selu.py:43
MASK_SHAPES = {"lstm2": (O.T, None, 2 * O.H), "l4": (None, O.L4_UNITS), "l5_1": (None, O.L5_UNITS), "l5_2": (None, O.L5_UNITS),
               "l5_3": (None, O.L5_UNITS), "l5_4": (None, O.L5_UNITS)}
DEFAULT_RATES = {"lstm2": 0.5, "l4": 0.5, "l5_1": 0.2, "l5_2": 0.2, "l5_3": 0.2, "l5_4": 0.2}     # clair/model.py:83-97


def make_masks(n, rates=DEFAULT_RATES, seed=0):
    """uint8 keep-masks (1 = kept) of one training batch, keyed like MASK_SHAPES."""
    rng = np.random.default_rng(seed)
    return {k: (rng.random(tuple(n if d is None else d for d in shp)) >= rates[k]).astype(np.uint8) for k, shp in MASK_SHAPES.items()}


def selu(x):
    return O.SELU_SCALE * torch.where(x >= 0, x, O.SELU_ALPHA * torch.expm1(torch.clamp(x, max=0.0)))


def dropout(x, mask, rate):
    return x if rate == 0 else x * mask / (1.0 - rate)


def alpha_dropout(x, mask, rate):
    if rate == 0:
        return x
    q = 1.0 - rate
    a = (1.0 / (q * ((1.0 - q) * ALPHA_DROPOUT ** 2 + 1.0))) ** 0.5
    b = -a * (1.0 - q) * ALPHA_DROPOUT
    return a * (x * mask + ALPHA_DROPOUT * (1.0 - mask)) + b


def lstm_direction(x_tm, kernel, bias, reverse):
    Tn, B, Fin = x_tm.shape
    Hn = kernel.shape[1] // 4
    xs = torch.flip(x_tm, [0]) if reverse else x_tm
    pre = (xs.reshape(Tn * B, Fin) @ kernel[:Fin] + bias).reshape(Tn, B, 4 * Hn)
    h = torch.zeros(B, Hn, dtype=x_tm.dtype)
    c = torch.zeros(B, Hn, dtype=x_tm.dtype)
    out = []
    for t in range(Tn):
        z = pre[t] + h @ kernel[Fin:]
        i, g, f, o = z[:, :Hn], z[:, Hn:2 * Hn], z[:, 2 * Hn:3 * Hn], z[:, 3 * Hn:]
        c = torch.tanh(g) * torch.sigmoid(i) + c * torch.sigmoid(f)
        h = torch.tanh(c) * torch.sigmoid(o)
        out.append(h)
    out = torch.stack(out)
    return torch.flip(out, [0]) if reverse else out


def bilstm(x_tm, w, layer):
    name = lambda d, v: O.LSTM_NAME.format(layer=layer, d=d, v=v)
    return torch.cat([lstm_direction(x_tm, w[name("fw", "kernel")], w[name("fw", "bias")], False),
                      lstm_direction(x_tm, w[name("bw", "kernel")], w[name("bw", "bias")], True)], dim=2)


def focal_loss(z, target):
    p = torch.softmax(z, dim=1)
    zeros = torch.zeros_like(p)
    pos = torch.where(target > 0, target - p, zeros)
    neg = torch.where(target > 0, zeros, p)
    return -((pos ** 2) * torch.log(torch.clamp(p, 1e-8, 1.0)) + (neg ** 2) * torch.log(torch.clamp(1.0 - p, 1e-8, 1.0))).sum()


def losses(X, Y, w, masks=None, rates=DEFAULT_RATES, l2_lambda=0.005, task_weights=(1, 1, 1, 1, 1)):
    """-> (total, [gt21, genotype, length 1, length 2, L2 without lambda]) as torch scalars; masks=None = inference phase."""
    train = masks is not None
    n = X.shape[0]
    m = {k: torch.from_numpy(np.asarray(v, dtype=np.float64)) for k, v in masks.items()} if train else {}
    r = rates if train else {k: 0.0 for k in rates}
    x_tm = X.reshape(n, O.T, O.F).transpose(0, 1)
    l1 = bilstm(x_tm, w, "LSTM1")
    l2 = bilstm(l1, w, "LSTM2")
    l2 = dropout(l2, m.get("lstm2"), r["lstm2"])
    k3 = torch.stack([w["L3/Unit_%d/kernel" % c] for c in range(2 * O.H)])
    b3 = torch.stack([w["L3/Unit_%d/bias" % c] for c in range(2 * O.H)])
    l3 = selu(torch.einsum("tbc,cto->boc", l2, k3) + b3.T[None])                        # [n,30,256]
    l4 = selu(l3.reshape(n, O.L3_UNITS * 2 * O.H) @ w["L4/kernel"] + w["L4/bias"])
    l4 = alpha_dropout(l4, m.get("l4"), r["l4"])
    parts, off = [], 0
    for k in range(4):
        a = selu(l4 @ w["L5_%d/kernel" % (k + 1)] + w["L5_%d/bias" % (k + 1)])
        a = alpha_dropout(a, m.get("l5_%d" % (k + 1)), r["l5_%d" % (k + 1)])
        z = selu(a @ w["Prediction/%s/kernel" % O.HEAD_NAMES[k]] + w["Prediction/%s/bias" % O.HEAD_NAMES[k]])
        parts.append(focal_loss(z, Y[:, off:off + O.HEADS[k]]))
        off += O.HEADS[k]
    parts.append(sum((v ** 2).sum() / 2 for name, v in w.items() if "bias" not in name))
    tw = task_weights
    total = tw[0] * parts[0] + tw[1] * parts[1] + tw[2] * parts[2] + tw[3] * parts[3] + tw[4] * l2_lambda * parts[4]
    return total, parts


def train_step(X, Y, weights, masks, rates=DEFAULT_RATES, l2_lambda=0.005, learning_rate=1e-3, clip_norm=5.0, adam_state=None,
               step=1):
    """One training step in float64 -> dict(loss, parts, grads, grad_norm, new_weights, adam_state)."""
    w = {k: torch.tensor(np.asarray(v, dtype=np.float64), requires_grad=True) for k, v in weights.items()}
    total, parts = losses(torch.from_numpy(np.asarray(X, dtype=np.float64)), torch.from_numpy(np.asarray(Y, dtype=np.float64)), w,
                          masks, rates, l2_lambda)
    total.backward()
    grads = {k: v.grad.numpy().copy() for k, v in w.items()}
    norm = float(np.sqrt(sum(float((g ** 2).sum()) for g in grads.values())))
    scale = clip_norm / max(norm, clip_norm)                                       # tf.clip_by_global_norm
    state = adam_state or {k: (np.zeros_like(g), np.zeros_like(g)) for k, g in grads.items()}
    b1, b2, eps = 0.9, 0.999, 1e-8
    lr_t = learning_rate * np.sqrt(1 - b2 ** step) / (1 - b1 ** step)
    new_w, new_state = {}, {}
    for k, g in grads.items():
        g = g * scale
        m_, v_ = state[k]
        m_ = b1 * m_ + (1 - b1) * g
        v_ = b2 * v_ + (1 - b2) * g * g
        new_w[k] = np.asarray(weights[k], dtype=np.float64) - lr_t * m_ / (np.sqrt(v_) + eps)
        new_state[k] = (m_, v_)
    return {"loss": float(total.detach()), "parts": [float(p.detach()) for p in parts], "grads": grads, "grad_norm": norm, "new_weights": new_w,
            "adam_state": new_state}
