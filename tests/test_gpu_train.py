"""Training step on the device against the float64 autograd restatement (oracle/train_oracle.py, whose
loss is pinned on the reference's own validate(): tests/test_train_oracle.py)."""
import os

import numpy as np
import pytest

from clair_b200 import synth
from oracle import train_oracle as TO

pytestmark = pytest.mark.gpu


def batch(n, seed):
    X = synth.synthetic_tensors(n, seed=seed)
    rng = np.random.default_rng(seed)
    Y = np.zeros((n, 90), np.float32)
    for lo, k in ((0, 21), (21, 3), (24, 33), (57, 33)):
        Y[np.arange(n), lo + rng.integers(0, k, n)] = 1
    return X, Y


def rel(a, b):
    return float(np.abs(a - b).max() / max(1e-30, np.abs(b).max()))


@pytest.mark.parametrize("n", [24, 21])                       # a multiple of the 8-site row tile, and a ragged batch
def test_losses_and_gradients_equal_autograd(weights1234, n):
    from clair_b200.train import Trainer
    X, Y = batch(n, 40 + n)
    masks = TO.make_masks(n, seed=7)
    want = TO.train_step(X, Y, weights1234, masks)
    t = Trainer(max_batch=64)
    t.set_weights(weights1234)
    parts = t.forward_backward(X, Y, masks)
    t.backward_lstm()
    np.testing.assert_allclose(parts, want["parts"], rtol=2e-5)
    assert abs(t.total_loss() - want["loss"]) <= 2e-5 * abs(want["loss"])
    got = t.gradients()
    # the oracle's gradients include the L2 term (lambda * w on kernels); the device adds it in apply()
    worst = {}
    for k, g in want["grads"].items():
        g_data = g - (0.005 * np.asarray(weights1234[k], np.float64) if "bias" not in k else 0.0)
        worst[k] = rel(got[k], g_data)
    bad = {k: v for k, v in worst.items() if v > 3e-4}
    assert not bad, sorted(bad.items(), key=lambda kv: -kv[1])[:5]
    # int16 transport of the same counts: identical step
    t2 = Trainer(max_batch=64)
    t2.set_weights(weights1234)
    parts16 = t2.forward_backward(X.astype(np.int16), Y, masks)
    np.testing.assert_allclose(parts16, parts, rtol=1e-6)         # (split-K sums are accumulated with atomics: order varies)
    t.close()
    t2.close()


def test_a_batch_beyond_the_machine_takes_the_64_site_backward_kernel(weights1234):
    """Up to ~590 sites both sequence kernels run 32 sites per cluster; beyond, the backward kernel runs 64 (train_engine.cuh:
    seq_rows_backward).  The loss and every gradient are sums over the sites of a batch, so the step on 620 sites (64-site
    kernel, ragged: ends inside a cluster) must equal the sum of the steps on its two halves (32-site kernel, the path the tests
    above hold against the oracle) - the float64 oracle itself needs a minute and a half for a batch of this size."""
    from clair_b200.train import Trainer
    n = 620
    X, Y = batch(n, 11)
    masks = TO.make_masks(n, seed=9)
    t = Trainer(max_batch=640)
    t.set_weights(weights1234)

    def run(lo, hi):
        m = {k: (v[:, lo:hi] if k == "lstm2" else v[lo:hi]) for k, v in masks.items()}
        parts = t.forward_backward(X[lo:hi], Y[lo:hi], m)
        t.backward_lstm()
        return np.array(parts[:4]), {k: v.astype(np.float64) for k, v in t.gradients().items()}

    whole_parts, whole = run(0, n)
    a_parts, a = run(0, n // 2)
    b_parts, b = run(n // 2, n)
    np.testing.assert_allclose(whole_parts, a_parts + b_parts, rtol=1e-5)
    worst = {k: rel(whole[k], a[k] + b[k]) for k in whole}
    bad = {k: v for k, v in worst.items() if v > 5e-5}
    assert not bad, sorted(bad.items(), key=lambda kv: -kv[1])[:5]
    t.close()


def test_two_adam_steps_equal_the_restatement(weights1234):
    from clair_b200.train import Trainer
    n = 16
    t = Trainer(max_batch=16)
    t.set_weights(weights1234)
    w, state = weights1234, None
    for step in (1, 2):
        X, Y = batch(n, 90 + step)
        masks = TO.make_masks(n, seed=step)
        want = TO.train_step(X, Y, w, masks, adam_state=state, step=step)
        loss = t.train(X, Y, masks)
        assert abs(loss - want["loss"]) <= 3e-5 * abs(want["loss"])
        assert abs(t.grad_norm - want["grad_norm"]) <= 1e-4 * want["grad_norm"] and t.grad_norm > 5.0     # the clip is active
        got = t.get_weights()
        for k in w:
            g = want["grads"][k]
            solid = np.abs(g) > 1e-3 * np.abs(g).max()             # where the sign-like first Adam steps are well defined
            d = np.abs(got[k] - want["new_weights"][k])
            assert d[solid].max() <= 1e-5, (k, step, d[solid].max())       # 1 % of the step size (lr = 1e-3)
            assert d.max() <= 2.1e-3                                # elsewhere at most the step size itself
        w, state = want["new_weights"], want["adam_state"]
        # Where a gradient is ~0 the first Adam step is +-lr by the sign of rounding noise, and such a weight then shifts every
        # gradient of step 2 a little: both sides take step 2 from the SAME weights (the device keeps its own Adam moments).
        t.set_weights({k: np.asarray(v, np.float32) for k, v in w.items()})
    t.close()


def test_device_drawn_masks_and_a_short_run(weights1234):
    from clair_b200.train import Trainer
    X, Y = batch(64, 5)
    a, b = Trainer(max_batch=64, seed=1), Trainer(max_batch=64, seed=1)
    c = Trainer(max_batch=64, seed=2)
    for t in (a, b, c):
        t.set_weights(weights1234)
    la, lb, lc = a.train(X, Y), b.train(X, Y), c.train(X, Y)
    assert abs(la - lb) <= 1e-6 * abs(la) and abs(la - lc) > 1e-4 * abs(la)      # the mask stream is a function of (seed, step)
    first = la
    for _ in range(14):
        last = a.train(X, Y)
    assert last < 0.7 * first                                       # fifteen steps on one batch: the loss falls
    # with every dropout off the training-phase probabilities are the inference probabilities of the same weights
    from clair_b200.model import Clair
    d = Trainer(max_batch=64, dropout_rates=(0, 0, 0, 0, 0, 0))
    d.set_weights(weights1234)
    d.forward_backward(X, Y)
    d.backward_lstm()
    m = Clair(max_sites=64)
    m.set_weights(weights1234)
    assert np.abs(d.probabilities() - m.predict_packed(X)).max() <= 1e-4
    # weights trained on the device load straight into the inference engine
    m.set_weights(a.get_weights())
    assert np.isfinite(m.predict_packed(X)).all()
    for t in (a, b, c, d):
        t.close()
    m.close()


def test_bad_arguments():
    from clair_b200.train import Trainer
    t = Trainer(max_batch=8)
    X, Y = batch(8, 1)
    with pytest.raises(ValueError, match="before the weights"):
        t.forward_backward(X, Y)
    with pytest.raises(ValueError):
        t.set_dropout_rates((1.0, 0, 0, 0, 0, 0))
    t.init()
    with pytest.raises(ValueError):
        t.forward_backward(np.zeros((9, 33, 8, 4), np.float32), np.zeros((9, 90), np.float32))
    with pytest.raises(ValueError, match="no forward_backward"):
        t.backward_lstm()
    t.forward_backward(X, Y)
    with pytest.raises(ValueError, match="backward_lstm has not run"):
        t.apply()
    t.close()


def _dp_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from clair_b200 import weights as W
    from clair_b200.train import DataParallelTrainer
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        w = W.random_weights(seed=1234)
        t = DataParallelTrainer(device=rank, max_batch=16)
        t.set_weights(w)
        X, Y = batch(32, 77)
        masks = TO.make_masks(32, seed=3)
        lo, hi = rank * 16, rank * 16 + 16
        mine = {k: (v[:, lo:hi] if k == "lstm2" else v[lo:hi]) for k, v in masks.items()}
        loss = t.train(X[lo:hi], Y[lo:hi], mine)
        q.put((rank, loss, t.grad_norm, {k: v for k, v in t.get_weights().items() if k in ("L4/bias", "L5_2/kernel")}))
        t.close()
    finally:
        dist.destroy_process_group()


def test_data_parallel_step_equals_the_single_gpu_step_on_the_whole_batch(weights1234):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from clair_b200.train import Trainer
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=300) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    X, Y = batch(32, 77)
    one = Trainer(max_batch=32)
    one.set_weights(weights1234)
    loss = one.train(X, Y, TO.make_masks(32, seed=3))
    ref = one.get_weights()
    for rank, l, norm, ws in results:
        assert abs(l - loss) <= 1e-5 * abs(loss)
        assert abs(norm - one.grad_norm) <= 1e-4 * one.grad_norm
        for k, v in ws.items():
            assert np.abs(v - ref[k]).max() <= 1e-5
    assert np.array_equal(results[0][3]["L4/bias"], results[1][3]["L4/bias"])      # every rank applied the same update
    one.close()
