"""The training-step oracle (oracle/train_oracle.py) against the reference's own loss graph."""
import os

import numpy as np
import torch

from oracle import train_oracle as TO

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_loss_restatement_equals_the_reference_validate(weights1234):
    # tests/golden/reference_model_loss.npz: what the REFERENCE's unmodified Clair.validate (clair/model.py:968-1008) returned
    # over the TensorFlow stand-in: total loss (lambda fed as 0), the four focal losses, the L2 term without lambda
    with np.load(os.path.join(GOLD, "reference_model_loss.npz")) as z:
        X, Y, want = z["X"].astype(np.float64), z["Y"].astype(np.float64), z["loss_f64"]
    w = {k: torch.tensor(np.asarray(v, dtype=np.float64)) for k, v in weights1234.items()}
    total, parts = TO.losses(torch.from_numpy(X), torch.from_numpy(Y), w, masks=None, l2_lambda=0.0)
    got = np.array([float(total)] + [float(p) for p in parts])
    np.testing.assert_allclose(got, want, rtol=1e-12)


def test_training_mode_pieces():
    # alpha-dropout keeps mean 0 / variance 1 of a standard normal input (that is its point: clair/selu.py:43-69)
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.standard_normal((200000,)))
    mask = torch.from_numpy((rng.random(200000) >= 0.5).astype(np.float64))
    y = TO.alpha_dropout(TO.selu(x) * 0 + x, mask, 0.5)
    assert abs(float(y.mean())) < 0.01 and abs(float(y.var()) - 1.0) < 0.02
    assert TO.alpha_dropout(x, mask, 0.0) is x and TO.dropout(x, mask, 0.0) is x
    np.testing.assert_allclose(TO.dropout(x, mask, 0.5).numpy(), (x * mask * 2).numpy())
    masks = TO.make_masks(7, seed=3)
    assert masks["lstm2"].shape == (33, 7, 256) and masks["l4"].shape == (7, 192) and masks["l5_3"].shape == (7, 96)


def test_one_step_moves_the_loss_down(weights1234):
    from clair_b200 import synth
    n = 12
    X = synth.synthetic_tensors(n, seed=8)
    rng = np.random.default_rng(1)
    Y = np.zeros((n, 90), np.float32)
    for lo, k in ((0, 21), (21, 3), (24, 33), (57, 33)):
        Y[np.arange(n), lo + rng.integers(0, k, n)] = 1
    masks = TO.make_masks(n, seed=5)
    r = TO.train_step(X, Y, weights1234, masks)
    assert r["grad_norm"] > 5.0                                    # so that the clip is exercised
    again = TO.train_step(X, Y, r["new_weights"], masks)
    assert again["loss"] < r["loss"]
    assert set(r["grads"]) == set(weights1234) and all(np.isfinite(g).all() for g in r["grads"].values())
