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


def test_three_term_tf32_split_is_fp32_grade():
    """The arithmetic of the training GEMMs (csrc/train_kernels.cuh: split_tf32 + mma_3xtf32), restated in numpy: an operand is
    hi (rounded to TF32's 10 mantissa bits in integer arithmetic) + lo (exact remainder, of which the tensor core reads the top
    19 bits), a product is a_lo.b_hi + a_hi.b_lo + a_hi.b_hi.  Its error stays below 2^-20 of |a.b| - the fp32 rounding of a
    single product is 2^-24 - and is not biased, which is what the near-cancelling sums of weight gradients need."""
    rng = np.random.default_rng(11)
    a = (rng.standard_normal(200000) * 10.0 ** rng.uniform(-6, 3, 200000)).astype(np.float32)
    b = (rng.standard_normal(200000) * 10.0 ** rng.uniform(-6, 3, 200000)).astype(np.float32)

    def split(x):
        bits = x.view(np.uint32)
        hi = ((bits + np.uint32(0x1000)) & np.uint32(0xffffe000)).view(np.float32)
        lo = (x - hi).astype(np.float32)                                  # exact in fp32
        assert np.array_equal(lo.astype(np.float64), x.astype(np.float64) - hi.astype(np.float64))
        lo_seen = (lo.view(np.uint32) & np.uint32(0xffffe000)).view(np.float32)      # what the MMA reads of the register
        return hi.astype(np.float64), lo_seen.astype(np.float64)

    ah, al = split(a)
    bh, bl = split(b)
    got = al * bh + ah * bl + ah * bh
    exact = a.astype(np.float64) * b.astype(np.float64)
    rel = (got - exact) / np.abs(exact)
    assert np.abs(rel).max() < 2.0 ** -20
    assert abs(rel.mean()) < 2.0 ** -26                                   # unbiased (measured 2^-33; worst single product 2^-21.1)
    # the rejected variant: hi cut instead of rounded biases every product the same way (mean error 2^-21.8 of the product)
    hi_cut = (a.view(np.uint32) & np.uint32(0xffffe000)).view(np.float32).astype(np.float64)
    lo_cut = ((a - hi_cut.astype(np.float32)).astype(np.float32).view(np.uint32) & np.uint32(0xffffe000)).view(np.float32).astype(np.float64)
    rel_cut = ((lo_cut * bh + hi_cut * bl + hi_cut * bh) - exact) / exact
    assert abs(rel_cut.mean()) > 10 * abs(rel.mean())
