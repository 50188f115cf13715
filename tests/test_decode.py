"""tensor_generator_from mirror vs fixtures produced by the REFERENCE's own generator
(oracle/gen_golden.py imports /root/reference/clair/utils.py to make them)."""
import gzip
import io
import os
import sys

import numpy as np

from clair_b200 import synth, utils

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_generator_matches_reference_fixture(capsys):
    with np.load(os.path.join(GOLDEN, "decode_expected.npz")) as z:
        exp = {k: z[k] for k in z.files}
    got = list(utils.tensor_generator_from(os.path.join(GOLDEN, "decode_rows.txt.gz"), 4))
    assert len(got) == 3
    total = 0
    for bi, (X, infos) in enumerate(got):
        assert X.dtype == np.float32
        np.testing.assert_array_equal(X, exp["X%d" % bi])          # bit-exact
        assert [" ".join(i) for i in infos] == list(exp["info%d" % bi])
        total += len(infos)
    assert total == 10                                             # the '*' centre row was dropped
    assert capsys.readouterr().err == str(exp["stderr"])           # same progress lines


def test_generator_reads_stdin_pipe(monkeypatch):
    counts = synth.synthetic_counts(5, seed=1)
    rows = [utils.format_tensor_row("chr1", 100 + i, "A" * 33, counts[i]) for i in range(5)]
    monkeypatch.setattr(sys, "stdin", io.StringIO("\n".join(rows) + "\n"))
    got = list(utils.tensor_generator_from("PIPE", 2))
    assert [len(i) for _, i in got] == [2, 2, 1]                   # ragged last batch
    X = np.concatenate([x for x, _ in got])
    np.testing.assert_array_equal(X, synth.synthetic_tensors(5, seed=1))


def test_all_rows_filtered_yields_nothing(tmp_path):
    counts = synth.synthetic_counts(3, seed=2)
    rows = [utils.format_tensor_row("c", i, "A" * 16 + "*" + "A" * 16, counts[i]) for i in range(3)]
    p = tmp_path / "t.gz"
    with gzip.open(p, "wt") as f:
        f.write("\n".join(rows) + "\n")
    assert list(utils.tensor_generator_from(str(p), 2)) == []


def test_empty_input_yields_nothing(tmp_path):
    p = tmp_path / "e.gz"
    with gzip.open(p, "wt") as f:
        f.write("")
    assert list(utils.tensor_generator_from(str(p), 4)) == []
