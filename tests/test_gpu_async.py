"""Many predict calls in flight (clairb_predict_async / clairb_predict_wait), the pipelined batch loop and the
multi-GPU entry of Clair on a real B200.  Everything here must be bit-identical to the synchronous single-GPU call: sites
are independent and every site goes through the same kernels whatever chunk, tile or device it lands on."""
import ctypes
import os
import threading

import numpy as np
import pytest

from clair_b200 import _lib, call_var, synth
from oracle import clair_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-4


def same(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)


def test_async_calls_equal_synchronous_calls_bit_for_bit(gpu_model):
    sizes = [1000, 1000, 333, 1, 2500, 1000, 7, 1000, 1000, 4096, 255, 257, 1000]
    Xs = [synth.synthetic_tensors(s, seed=900 + i) for i, s in enumerate(sizes)]
    want = [[a.copy() for a in gpu_model.predict(X)] for X in Xs]
    tickets = [gpu_model.predict_async(X) for X in Xs]          # all in flight before the first wait
    got = [t.result() for t in tickets]
    for g, w, s in zip(got, want, sizes):
        assert [a.shape for a in g] == [(s, 21), (s, 3), (s, 33), (s, 33)]
        same(g, w)
    assert gpu_model.prediction is got[-1]
    # against the oracle too, so that both being wrong together is excluded
    ref = O.forward(Xs[2], gpu_model_weights(gpu_model), np.float64)
    for k in range(4):
        assert np.abs(got[2][k] - ref[k]).max() <= TOL


_W = {}


def gpu_model_weights(m):
    from clair_b200 import weights as W
    if "w" not in _W:
        _W["w"] = W.random_weights(seed=1234)
    return _W["w"]


def test_async_waits_in_any_order_and_requests_larger_than_a_chunk(weights1234):
    from clair_b200.model import Clair
    m = Clair(max_sites=4096, batch_sites=1000)                 # chunk = 4096 sites: a 10,000-site request spans three
    m.set_weights(weights1234)
    big = synth.synthetic_tensors(10000, seed=31)
    small = [synth.synthetic_tensors(s, seed=32 + s) for s in (5, 1000, 3000)]
    t_big = m.predict_async(big)
    t_small = [m.predict_async(X) for X in small]
    got_small = [t.result() for t in reversed(t_small)][::-1]   # later tickets waited for first
    got_big = t_big.result()
    want_big = np.concatenate([m.predict_packed(big[i:i + 4000]) for i in range(0, 10000, 4000)])
    np.testing.assert_array_equal(np.concatenate(got_big, axis=1), want_big)
    for X, g in zip(small, got_small):
        np.testing.assert_array_equal(np.concatenate(g, axis=1), m.predict_packed(X))
    m.close()


def test_async_mixed_dtypes_layouts_and_decisions_share_the_queue(gpu_model):
    from clair_b200 import decision
    lib = _lib.load()
    rng = np.random.default_rng(3)
    jobs = []
    for i in range(24):
        n = int(rng.integers(1, 1500))
        X = synth.synthetic_tensors(n, seed=1200 + i)
        kind = i % 4
        if kind == 1:
            X = X.astype(np.int16)
        jobs.append((kind, X, (np.arange(n) % 4).astype(np.uint8)))
    tickets, packed_outs = [], []
    for kind, X, ref in jobs:
        if kind == 2:                                           # packed rows straight through the C-ABI
            out = np.empty((len(X), 90), np.float32)
            t = ctypes.c_int64()
            rc = lib.clairb_predict_async(gpu_model._h, X.ctypes.data_as(ctypes.c_void_p), _lib.DTYPE_F32, len(X),
                                          out.ctypes.data_as(ctypes.c_void_p), None, None, None, None, None, ctypes.byref(t))
            assert rc == 0
            tickets.append(t.value)
            packed_outs.append(out)
        elif kind == 3:
            tickets.append(gpu_model.predict_async(X, ref))
        else:
            tickets.append(gpu_model.predict_async(X))
    packed_outs = iter(packed_outs)
    for (kind, X, ref), t in zip(jobs, tickets):
        if kind == 2:
            assert lib.clairb_predict_wait(gpu_model._h, t) == 0
            np.testing.assert_array_equal(next(packed_outs), gpu_model.predict_packed(X))
        elif kind == 3:
            pred, dec = t.result()
            want_pred, want_dec = gpu_model.predict_and_decide(X, ref)
            same(pred, want_pred)
            for f in decision.Decision._fields:
                np.testing.assert_array_equal(getattr(dec, f), getattr(want_dec, f))
        else:
            same(t.result(), gpu_model.predict(X))


def test_async_bad_arguments_and_unknown_tickets(gpu_model):
    lib = _lib.load()
    X = synth.synthetic_tensors(4, seed=1)
    out = np.empty((4, 90), np.float32)
    t = ctypes.c_int64()
    xp, op = X.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p)
    assert lib.clairb_predict_async(gpu_model._h, xp, 0, 0, op, None, None, None, None, None, ctypes.byref(t)) == _lib.EINVAL
    assert lib.clairb_predict_async(gpu_model._h, xp, 0, 4, op, op, None, None, None, None, ctypes.byref(t)) == _lib.EINVAL
    assert lib.clairb_predict_async(gpu_model._h, xp, 7, 4, op, None, None, None, None, None, ctypes.byref(t)) == _lib.EINVAL
    assert lib.clairb_predict_async(gpu_model._h, xp, 0, 4, op, None, None, None, xp, None, ctypes.byref(t)) == _lib.EINVAL
    assert lib.clairb_predict_wait(gpu_model._h, 987654321) == _lib.EINVAL
    assert "ticket" in _lib.last_error(gpu_model._h)
    with pytest.raises(ValueError):
        gpu_model.predict_async(np.zeros((3, 33, 8, 5), np.float32))


def test_async_submitters_on_several_threads_and_a_synchronous_call_in_between(gpu_model):
    Xs = [synth.synthetic_tensors(600 + 13 * i, seed=70 + i) for i in range(16)]
    want = [np.concatenate(gpu_model.predict(X), axis=1) for X in Xs]
    got = [None] * len(Xs)

    def submitter(lo, hi):
        tickets = [(i, gpu_model.predict_async(Xs[i])) for i in range(lo, hi)]
        for i, t in tickets:
            got[i] = np.concatenate(t.result(), axis=1)

    threads = [threading.Thread(target=submitter, args=(i, i + 4)) for i in range(0, 16, 4)]
    for t in threads:
        t.start()
    mid = gpu_model.predict_packed(Xs[5])                       # serialised behind whatever is queued
    for t in threads:
        t.join()
    np.testing.assert_array_equal(mid, want[5])
    for g, w in zip(got, want):
        np.testing.assert_array_equal(g, w)


def test_destroy_with_unwaited_tickets(weights1234):
    from clair_b200.model import Clair
    m = Clair(max_sites=2048)
    m.set_weights(weights1234)
    X = synth.synthetic_tensors(1500, seed=9)
    tickets = [m.predict_async(X) for _ in range(6)]
    first = tickets[0].result()
    m.close()                                                   # completes what is queued, then tears down
    del tickets
    assert np.isfinite(first[0]).all()


def test_pipelined_batch_loop_on_the_device(gpu_model, weights1234):
    from clair_b200.model import PinnedPool
    sizes = [1000] * 45 + [333]
    pool = PinnedPool(24, (1000, 1056), np.float32)
    taken = []

    def gen():
        for i, s in enumerate(sizes):
            buf = pool.take((1000, 1056), np.float32)
            X = buf.reshape(1000, 33, 8, 4)[:s]
            X[...] = synth.synthetic_tensors(s, seed=5000 + i)
            taken.append(i)
            yield X, [["c", str(1000 * i + j), "A" * 33] for j in range(s)]

    got = []

    def output(mb, Y):
        got.append((int(mb[1][0][1]) // 1000, [y.copy() for y in Y], mb[0].copy()))

    call_var.run_batches(gpu_model, gen(), output, in_flight=8, release=pool.give)
    pool.close()
    assert [i for i, _, _ in got] == list(range(len(sizes)))
    for i, Y, X in got:
        np.testing.assert_array_equal(X, synth.synthetic_tensors(sizes[i], seed=5000 + i))    # buffers were not recycled early
        same(Y, gpu_model.predict(X))
    ref = O.forward(got[-1][2], weights1234, np.float64)
    for k in range(4):
        assert np.abs(got[-1][1][k] - ref[k]).max() <= TOL
        np.testing.assert_array_equal(got[-1][1][k].argmax(1), ref[k].argmax(1))


def test_pipelined_batch_loop_with_decision_records(gpu_model):
    from clair_b200 import decision
    sizes = [1000, 1000, 500]
    items = []
    for i, s in enumerate(sizes):
        X = synth.synthetic_tensors(s, seed=40 + i)
        items.append((X, [["chr1", str(j), "A" * 16 + "ACGT"[(i + j) % 4] + "A" * 16] for j in range(s)]))
    got = []
    call_var.run_batches(gpu_model, iter(items), lambda mb, Y, dec: got.append((Y, dec)), with_decision=True, in_flight=4)
    for (X, infos), (Y, dec) in zip(items, got):
        want_Y, want_dec = gpu_model.predict_and_decide(X, decision.ref_base_codes(infos))
        same(Y, want_Y)
        for f in decision.Decision._fields:
            np.testing.assert_array_equal(getattr(dec, f), getattr(want_dec, f))


def test_predict_to_device_leaves_the_packed_rows_in_device_memory(gpu_model):
    import torch
    X = synth.synthetic_tensors(3000, seed=77)
    od = torch.zeros((3000, 90), dtype=torch.float32, device="cuda")
    rc = _lib.load().clairb_predict_to_device(gpu_model._h, X.ctypes.data_as(ctypes.c_void_p), _lib.DTYPE_F32, 3000,
                                              ctypes.c_void_p(od.data_ptr()))
    assert rc == 0
    np.testing.assert_array_equal(od.cpu().numpy(), gpu_model.predict_packed(X))


def _device_count():
    import torch
    return torch.cuda.device_count()


def test_multi_device_predict_equals_single_device(gpu_model, weights1234):
    if _device_count() < 2:
        pytest.skip("needs at least two GPUs (gpurun --gpus 2)")
    from clair_b200.model import Clair
    g = min(_device_count(), 8)
    m = Clair(devices=list(range(g)), max_sites=65536)
    m.set_weights(weights1234)
    for n in (g * 4096 + 37, 5, 1000):
        X = synth.synthetic_tensors(n, seed=n)
        same(m.predict(X), gpu_model.predict(X) if n <= 8192 else
             [np.concatenate(p) for p in zip(*[gpu_model.predict(X[i:i + 8192]) for i in range(0, n, 8192)])])
    # batches of the loop go round-robin to the GPUs, results still in order
    Xs = [synth.synthetic_tensors(1000, seed=300 + i) for i in range(3 * g)]
    tickets = [m.predict_async(X) for X in Xs]
    for X, t in zip(Xs, tickets):
        same(t.result(), gpu_model.predict(X))
    m.close()


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="needs a GPU and the reference tree on the same machine")
def test_reference_call_variants_unmodified_on_the_device(gpu_model, weights1234, tmp_path):
    """SURVEY.md 8b acceptance on the real library: see tests/test_async_loop.py for the loop and what it checks; here the
    Clair under the reference's call_variants is the device one."""
    import gzip
    import types
    from clair_b200 import utils
    from oracle import gen_golden_decision as GD
    cv = GD.import_reference_call_var()
    n = 2350
    counts = synth.synthetic_counts(n, seed=77)
    rng = np.random.default_rng(5)
    path = tmp_path / "tensors.gz"
    with gzip.open(path, "wt") as f:
        for i in range(n):
            f.write(utils.format_tensor_row("chr20", 1000 + 50 * i, "".join(rng.choice(list("ACGT"), size=33)), counts[i]) + "\n")
    config = cv.OutputConfig(is_show_reference=True, is_debug=False, is_haploid_precision_mode_enabled=False,
                             is_haploid_sensitive_mode_enabled=False, is_output_for_ensemble=False, quality_score_for_pass=None)
    lines, rec = [], GD.Recorder()
    util = cv.OutputUtilities(print_debug_message=lambda *a: None, insertion_bases_using=rec.insertion_bases_using,
                              deletion_bases_using=rec.deletion_bases_using,
                              insertion_bases_using_pysam_using=rec.insertion_bases_using_pysam_using, output=lines.append,
                              output_header=lambda: None, close_opened_files=lambda: None)
    cv.call_variants(types.SimpleNamespace(tensor_fn=str(path)), gpu_model, config, util)
    assert len(lines) == n


def test_queued_calls_with_device_output(gpu_model):
    import torch
    lib = _lib.load()
    X = synth.synthetic_tensors(5000, seed=123).astype(np.int16)
    od = torch.zeros((5000, 90), dtype=torch.float32, device="cuda")
    bounds = [(0, 700), (700, 3300), (3300, 5000)]
    tickets = []
    for lo, hi in bounds:
        t = ctypes.c_int64()
        rc = lib.clairb_predict_async_to_device(gpu_model._h, X[lo:hi].ctypes.data_as(ctypes.c_void_p), _lib.DTYPE_I16, hi - lo,
                                                ctypes.c_void_p(od[lo:hi].data_ptr()), ctypes.byref(t))
        assert rc == 0
        tickets.append(t.value)
    mixed = gpu_model.predict_async(X[:100])                     # a host-output call in the same queue
    for t in tickets:
        assert lib.clairb_predict_wait(gpu_model._h, t) == 0
    np.testing.assert_array_equal(od.cpu().numpy(), gpu_model.predict_packed(X))
    np.testing.assert_array_equal(np.concatenate(mixed.result(), axis=1), gpu_model.predict_packed(X[:100]))
    t = ctypes.c_int64()
    assert lib.clairb_predict_async_to_device(gpu_model._h, X.ctypes.data_as(ctypes.c_void_p), 1, 10, None, ctypes.byref(t)) == _lib.EINVAL
