"""The pipelined batch loop (clair_b200.call_var.run_batches with in_flight > 1) on the host, and the reference's own
unmodified call_variants loop (clair/call_var.py:1312-1367) driven against clair_b200.model.Clair.

No GPU here: the model under the loops is either a fake with the predict_async surface or the real Clair class over a
stand-in for the C-ABI library that answers from the CPU oracle (test infrastructure; the product path has no such
fallback - see tests/test_gpu_async.py for the same loops on the device).
"""
import ctypes
import os
import threading
import time

import numpy as np
import pytest

from clair_b200 import call_var, synth


class FakeTicket:
    def __init__(self, model, X, delay):
        self.model, self.X, self.delay = model, X, delay

    def result(self):
        if self.delay:
            time.sleep(self.delay)
        self.model.in_flight -= 1
        pred = [self.X.reshape(self.X.shape[0], -1).sum(1)]
        self.model.prediction = pred
        return pred


class FakeAsyncModel:
    def __init__(self, delay=0.0):
        self.prediction, self.delay = None, delay
        self.in_flight = self.max_in_flight = 0
        self.submitted = []

    def predict_async(self, batchX, ref_bases=None):
        self.in_flight += 1
        self.max_in_flight = max(self.max_in_flight, self.in_flight)
        self.submitted.append(batchX.shape[0])
        return FakeTicket(self, batchX, self.delay)


def batches(sizes, seed=100):
    off = 0
    for s in sizes:
        X = synth.synthetic_tensors(s, seed=seed + off)
        yield X, [["chr1", str(off + i), "A" * 33] for i in range(s)]
        off += s


def test_pipelined_loop_keeps_order_and_pairs_batches_with_their_predictions():
    m = FakeAsyncModel(delay=0.002)
    seen, released = [], []

    def output(mini_batch, batch_Y, tag):
        X, infos = mini_batch
        np.testing.assert_array_equal(batch_Y[0], X.reshape(X.shape[0], -1).sum(1))
        seen.append((len(infos), infos[0][1], tag))

    sizes = [5, 5, 5, 5, 5, 5, 5, 5, 5, 3]
    call_var.run_batches(m, batches(sizes), output, "cfg", in_flight=4, release=lambda X: released.append(X.shape[0]))
    assert [s for s, _, _ in seen] == sizes
    assert [p for _, p, _ in seen] == [str(5 * i) for i in range(10)]
    assert m.submitted == sizes and released == sizes
    assert 1 < m.max_in_flight <= 4                # several in flight, never more than asked for


def test_pipelined_loop_is_the_default_for_async_models_and_handles_an_empty_source():
    m = FakeAsyncModel()
    got = []
    call_var.run_batches(m, batches([4, 4, 2]), lambda mb, Y: got.append(len(mb[1])))
    assert got == [4, 4, 2]
    call_var.run_batches(m, iter(()), lambda *a: pytest.fail("no output expected"))


def test_pipelined_loop_propagates_failures_from_every_stage_and_still_waits_for_all_tickets():
    def bad_source():
        yield from batches([3, 3])
        raise OSError("tensor stream broke")

    m = FakeAsyncModel()
    with pytest.raises(OSError, match="tensor stream broke"):
        call_var.run_batches(m, bad_source(), lambda mb, Y: None, in_flight=4)
    assert m.in_flight == 0

    m = FakeAsyncModel()

    def bad_output(mb, Y):
        if mb[1][0][1] == "3":
            raise KeyError("vcf writer")

    with pytest.raises(KeyError):
        call_var.run_batches(m, batches([3, 3, 3, 3, 3]), bad_output, in_flight=2)
    assert m.in_flight == 0                        # the tickets behind the failure were still collected

    class Refuses(FakeAsyncModel):
        def predict_async(self, batchX, ref_bases=None):
            if len(self.submitted) == 2:
                raise ValueError("Inconsistent shape")
            return FakeAsyncModel.predict_async(self, batchX, ref_bases)

    m = Refuses()
    with pytest.raises(ValueError, match="Inconsistent shape"):
        call_var.run_batches(m, batches([3, 3, 3, 3]), lambda mb, Y: None, in_flight=3)
    assert m.in_flight == 0


def test_lock_step_loop_needs_no_async_surface_and_in_flight_above_one_does():
    class SyncOnly:
        prediction = None

        def predict(self, batchX):
            self.prediction = [batchX.shape[0]]

    got = []
    call_var.run_batches(SyncOnly(), batches([2, 2]), lambda mb, Y: got.append(Y[0]))
    assert got == [2, 2]
    with pytest.raises(ValueError):
        call_var.run_batches(SyncOnly(), batches([2]), lambda mb, Y: None, in_flight=8)


# ---- the reference's own loop against the Clair class ------------------------------------------------------------------
class OracleLib:
    """Stand-in for libclair_b200.so behind clair_b200.model.Clair: the entry points the class calls on the reference's
    loop, answered by the numpy oracle.  TEST INFRASTRUCTURE ONLY (lets the host side of the drop-in run where there is
    no GPU); it is injected by the test and is not reachable from the package."""

    def __init__(self):
        self.weights, self.calls, self.threads = {}, [], set()

    def clairb_create(self, device, max_sites, batch_sites, out):
        out._obj.value = 0x1234
        return 0

    def clairb_set_weight(self, h, name, data, shape, rank):
        dims = [shape[i] for i in range(rank)]
        count = int(np.prod(dims))
        src = (ctypes.c_float * count).from_address(data.value)
        self.weights[name.decode()] = np.ctypeslib.as_array(src).reshape(dims).copy()
        return 0

    def clairb_finalize_weights(self, h):
        return 0

    def clairb_predict_split(self, h, x, dtype, n, *outs):
        from oracle import clair_oracle as O
        self.threads.add(threading.current_thread().name)
        X = np.ctypeslib.as_array((ctypes.c_float * (n * 1056)).from_address(x.value)).reshape(n, 33, 8, 4)
        self.calls.append(n)
        probs = O.forward(X, self.weights, np.float32)
        for ptr, p in zip(outs, probs):
            dst = np.ctypeslib.as_array((ctypes.c_float * p.size).from_address(ptr.value)).reshape(p.shape)
            dst[...] = p
        return 0

    def clairb_destroy(self, h):
        return 0

    def clairb_last_error(self, h):
        return b""


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="runs the reference's own call_variants; only where /root/reference exists")
def test_reference_call_variants_runs_unmodified_against_clair_b200_model(monkeypatch, weights1234, tmp_path):
    """SURVEY.md 8b acceptance: `call_variants(args, m, output_config, output_utilities)` of the reference, imported from
    /root/reference and not touched, with m = clair_b200.model.Clair.  Reference generator (gzip -fdc child), reference
    threads (predict runs in a threading.Thread, kwargs batchX=), reference batch_output reading m.prediction one iteration
    later, reference VCF rows.  The rows must be the ones the same loop prints over a model that answers straight from the
    oracle."""
    import gzip
    import types
    from clair_b200 import _lib, utils
    from clair_b200.model import Clair
    from oracle import clair_oracle as O
    from oracle import gen_golden_decision as GD
    cv = GD.import_reference_call_var()

    # 2,350 rows of CreateTensor text (three batches of param.predictBatchSize = 1000, the last one ragged)
    n = 2350
    counts = synth.synthetic_counts(n, seed=77)
    rng = np.random.default_rng(5)
    path = tmp_path / "tensors.gz"
    with gzip.open(path, "wt") as f:
        for i in range(n):
            seq = "".join(rng.choice(list("ACGT"), size=33))
            f.write(utils.format_tensor_row("chr20", 1000 + 50 * i, seq, counts[i]) + "\n")
    args = types.SimpleNamespace(tensor_fn=str(path))
    config = cv.OutputConfig(is_show_reference=True, is_debug=False, is_haploid_precision_mode_enabled=False,
                             is_haploid_sensitive_mode_enabled=False, is_output_for_ensemble=False, quality_score_for_pass=None)

    def run(m):
        lines, rec, marks = [], GD.Recorder(), []
        util = cv.OutputUtilities(print_debug_message=lambda *a: None, insertion_bases_using=rec.insertion_bases_using,
                                  deletion_bases_using=rec.deletion_bases_using,
                                  insertion_bases_using_pysam_using=rec.insertion_bases_using_pysam_using, output=lines.append,
                                  output_header=lambda: marks.append("header"), close_opened_files=lambda: marks.append("closed"))
        cv.call_variants(args, m, config, util)
        assert marks == ["header", "closed"]
        return lines

    class StraightFromOracle:
        prediction = None

        def predict(self, batchX):
            self.prediction = O.forward(np.asarray(batchX, np.float32), weights1234, np.float32)

    want = run(StraightFromOracle())

    lib = OracleLib()
    monkeypatch.setattr(_lib, "load", lambda path=None: lib)
    m = Clair()                                   # reference: Clair() then restore_parameters / init
    m.set_weights(weights1234)
    got = run(m)
    m.close()
    assert lib.calls == [1000, 1000, 350]
    assert "MainThread" not in lib.threads         # the reference runs predict in its own Thread per iteration
    assert len(want) == n and got == want
    assert [a.shape for a in m.prediction] == [(350, 21), (350, 3), (350, 33), (350, 33)]
