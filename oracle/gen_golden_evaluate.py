"""Golden report of the reference's own evaluate_model (clair/evaluate.py:37-148), the second caller of Clair.predict.
TEST INFRASTRUCTURE.  Run from the repo root (only where /root/reference exists):  python oracle/gen_golden_evaluate.py

Imports /root/reference/clair/evaluate.py with its absent third-party imports stubbed (tensorflow via clair.model, blosc ->
clair_b200.bins.unpack_array, intervaltree) and `np.int` restored (removed from numpy 1.24; the reference was written for
1.x), runs evaluate_model(m, dataset_info) on a seeded bin with `m` = FakeModel below (integer arithmetic, no ties, so its
arg-max and top-2 order are the same on every machine), and stores what it prints: tests/golden/evaluate_report.txt.
"""
import contextlib
import io
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REFERENCE = "/root/reference"
N_SITES = 1730                       # 3 full frames of 500 + one of 230; batches of 1000 straddle frames


class FakeModel(object):
    """predict(x) -> four probability arrays computed with exact integer arithmetic from the tensors."""

    def __init__(self, seed=77):
        self.W = np.random.default_rng(seed).integers(-7, 8, size=(1056, 90)).astype(np.int64)
        self.calls = []

    def predict(self, batchX):
        x = np.asarray(batchX).reshape(len(batchX), -1).astype(np.int64)
        self.calls.append(len(x))
        score = (x @ self.W) % 997 * 100 + np.arange(90)           # distinct within a row: no ties anywhere
        out = []
        for a, b in ((0, 21), (21, 24), (24, 57), (57, 90)):
            s = score[:, a:b].astype(np.float64)
            out.append((s / s.sum(axis=1, keepdims=True)).astype(np.float32))
        self.prediction = out
        return out


def make_dataset(n=N_SITES, seed=99):
    from clair_b200 import synth
    rng = np.random.default_rng(seed)
    x = synth.synthetic_tensors(n, seed=seed)
    y = np.zeros((n, 90), np.float32)
    for a, b in ((0, 21), (21, 24), (24, 57), (57, 90)):
        y[np.arange(n), a + rng.integers(0, b - a, size=n)] = 1
    pos = np.array(["chr1:%d" % (1000 + 3 * i) for i in range(n)])
    return x, y, pos


def dataset_info(tmp_dir):
    from clair_b200 import bins
    x, y, pos = make_dataset()
    path = os.path.join(tmp_dir, "evaluate.bin")
    bins.write_bin(path, x, y, pos)
    return bins.dataset_info_from(binary_file_path=path)


def report_of(evaluate_model, info):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        evaluate_model(FakeModel(), info)
    return buf.getvalue()


def main():
    if not os.path.isdir(REFERENCE):
        print("no /root/reference here: evaluate_report.txt not regenerated")
        return
    import tempfile
    from clair_b200 import bins
    sys.path.insert(0, REFERENCE)
    stub = types.ModuleType("blosc")
    stub.unpack_array = bins.unpack_array
    stub.pack_array = lambda a, **kw: bins.pack_array(a)
    stub.set_nthreads = lambda n: None
    stub.NOSHUFFLE = 0
    sys.modules["blosc"] = stub
    tree = types.ModuleType("intervaltree")
    tree.IntervalTree = object
    sys.modules.setdefault("intervaltree", tree)
    import clair
    model_stub = types.ModuleType("clair.model")
    model_stub.Clair = object
    sys.modules["clair.model"] = model_stub
    clair.model = model_stub
    if not hasattr(np, "int"):
        np.int = int                                     # numpy < 1.24 spelling used at clair/evaluate.py:34
    import clair.evaluate as ref_evaluate
    with tempfile.TemporaryDirectory() as tmp:
        text = report_of(ref_evaluate.evaluate_model, dataset_info(tmp))
    path = os.path.join(ROOT, "tests", "golden", "evaluate_report.txt")
    with open(path, "w") as f:
        f.write(text)
    print(path, "written:", len(text.splitlines()), "lines;", text.splitlines()[1])


if __name__ == "__main__":
    main()
