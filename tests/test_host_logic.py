"""Host-side logic that needs no GPU: batch loop contract, weight blob, sharding (+ gloo x2)."""
import os
import sys
import threading

import numpy as np
import pytest

from clair_b200 import call_var, shard, synth, weights as W


class FakeModel:
    """Stands in for Clair: prediction = per-site checksum, records call order and threads."""

    def __init__(self):
        self.prediction = None
        self.calls = []
        self.in_flight = 0
        self.max_in_flight = 0

    def predict(self, batchX):
        self.in_flight += 1
        self.max_in_flight = max(self.max_in_flight, self.in_flight)
        self.calls.append((threading.current_thread() is threading.main_thread(), batchX.shape[0]))
        self.prediction = [batchX.reshape(batchX.shape[0], -1).sum(1)]
        self.in_flight -= 1
        return self.prediction


def batches(sizes):
    off = 0
    for s in sizes:
        X = synth.synthetic_tensors(s, seed=100 + off)
        yield X, [["chr1", str(off + i), "A" * 33] for i in range(s)]
        off += s


def test_run_batches_keeps_order_and_hands_over_previous_prediction():
    m = FakeModel()
    seen = []

    def output(mini_batch, batch_Y, tag):
        X, infos = mini_batch
        # the prediction handed over must belong to exactly this batch (call_var.py:1334-1338)
        np.testing.assert_array_equal(batch_Y[0], X.reshape(X.shape[0], -1).sum(1))
        seen.append((len(infos), infos[0][1], tag))

    sizes = [4, 4, 4, 3]                                    # ragged last batch
    call_var.run_batches(m, batches(sizes), output, "cfg")
    assert [s for s, _, _ in seen] == sizes
    assert [p for _, p, _ in seen] == ["0", "4", "8", "12"]
    assert [n for _, n in m.calls] == sizes
    assert not any(is_main for is_main, _ in m.calls)       # predict runs off the main thread
    assert m.max_in_flight == 1


def test_run_batches_empty_source():
    m = FakeModel()
    call_var.run_batches(m, iter(()), lambda *a: pytest.fail("no output expected"))
    assert m.calls == []


def test_weight_blob_round_trip(tmp_path, weights1234):
    p = tmp_path / "model"
    W.save_blob(str(p) + ".npz", weights1234)
    w = W.load_blob(str(p))                                  # suffix added like a checkpoint prefix
    assert set(w) == set(weights1234)
    for k in w:
        np.testing.assert_array_equal(w[k], weights1234[k])
    bad = dict(weights1234)
    bad.pop("L4/bias")
    with pytest.raises(ValueError):
        W.check_weights(bad)
    bad = dict(weights1234)
    bad["L4/bias"] = np.zeros(191, np.float32)
    with pytest.raises(ValueError):
        W.check_weights(bad)


def test_synthetic_tensors_look_like_generator_output():
    c = synth.synthetic_counts(64, seed=5)
    x = synth.synthetic_tensors(64, seed=5)
    assert c.dtype == np.int16 and c.min() >= 0
    assert x.dtype == np.float32 and x.shape == (64, 33, 8, 4)
    assert (x == np.round(x)).all() and x.min() < 0          # small signed integers
    np.testing.assert_array_equal(x[..., 0], c[..., 0])
    np.testing.assert_array_equal(x[..., 2], c[..., 2] - c[..., 0])


@pytest.mark.parametrize("n,g", [(1000, 8), (7, 8), (1, 2), (4096, 3), (0, 4)])
def test_shard_bounds_partition(n, g):
    parts = [shard.shard_bounds(n, g, r) for r in range(g)]
    assert parts[0][0] == 0 and parts[-1][1] == n
    for (a, b), (c, d) in zip(parts, parts[1:]):
        assert b == c and a <= b and c <= d


def _gloo_worker(rank, world, port, n, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        X = synth.synthetic_tensors(n, seed=42)
        fake = lambda x: np.tile(x.reshape(x.shape[0], -1)[:, :90] * 0.5 + 1.0, 1).astype(np.float32)
        full = shard.sharded_predict(fake, X, rank, world)
        if rank == 0:
            q.put(np.array_equal(full, fake(X)))
        else:
            q.put(full is None)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [37, 2])
def test_sharded_predict_two_ranks_gloo(n):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n) % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(results)


def _gradient_exchange_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from clair_b200.train import GradientExchange
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, dense_offset = 1000, 377
        mine = torch.arange(n, dtype=torch.float32) * (rank + 1) + rank
        want = sum(torch.arange(n, dtype=torch.float32) * (r + 1) + r for r in range(world))
        ex = GradientExchange(mine, dense_offset, dist)
        ex.dense()                                             # the tail first ...
        ex.wait()
        tail_done = torch.equal(mine[dense_offset:], want[dense_offset:]) and not torch.equal(mine[:dense_offset], want[:dense_offset])
        ex.head()                                              # ... then the LSTM part
        ex.wait()
        q.put(bool(tail_done and torch.equal(mine, want)))
    finally:
        dist.destroy_process_group()


def test_gradient_exchange_two_ranks_gloo():
    """The data-parallel trainer's all-reduce of the flat gradient buffer (SUM, two pieces) on world_size 2 over gloo."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gradient_exchange_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(results)


def test_run_batches_with_decision_hands_over_matching_records():
    from clair_b200 import decision

    class FakeDecider(FakeModel):
        decision = None

        def predict_and_decide(self, batchX, ref_bases):
            self.predict(batchX)
            n = batchX.shape[0]
            first = int(batchX.reshape(n, -1).sum(1)[0])
            return self.prediction, decision.Decision(np.full(n, first), np.asarray(ref_bases), None, None, None, None)

    m = FakeDecider()
    seen = []

    def output(mini_batch, batch_Y, dec, tag):
        X, infos = mini_batch
        assert dec.category[0] == int(X.reshape(X.shape[0], -1).sum(1)[0])      # records of exactly this batch
        assert dec.len1.tolist() == ["ACGT".index(i[2][16]) for i in infos]       # reference bases came from the infos
        seen.append((len(infos), tag))

    def gen():
        off = 0
        for s in (3, 3, 2):
            X = synth.synthetic_tensors(s, seed=300 + off)
            yield X, [["chr1", str(off + i), "A" * 16 + "ACGT"[(off + i) % 4] + "A" * 16] for i in range(s)]
            off += s

    call_var.run_batches(m, gen(), output, "cfg", with_decision=True)
    assert seen == [(3, "cfg"), (3, "cfg"), (2, "cfg")]
    assert m.max_in_flight == 1
