"""tensor_generator_from mirror vs fixtures produced by the REFERENCE's own generator
(oracle/gen_golden.py imports /root/reference/clair/utils.py to make them)."""
import gzip
import io
import os
import sys

import numpy as np

from clair_b200 import synth, utils

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


import pytest


@pytest.mark.parametrize("decoder", ["native", "python"])
def test_generator_matches_reference_fixture(capsys, decoder):
    with np.load(os.path.join(GOLDEN, "decode_expected.npz")) as z:
        exp = {k: z[k] for k in z.files}
    got = list(utils.tensor_generator_from(os.path.join(GOLDEN, "decode_rows.txt.gz"), 4, decoder=decoder))
    assert len(got) == 3
    total = 0
    for bi, (X, infos) in enumerate(got):
        assert X.dtype == np.float32
        np.testing.assert_array_equal(X, exp["X%d" % bi])          # bit-exact
        assert [" ".join(i) for i in infos] == list(exp["info%d" % bi])
        total += len(infos)
    assert total == 10                                             # the '*' centre row was dropped
    assert capsys.readouterr().err == str(exp["stderr"])           # same progress lines


@pytest.mark.parametrize("decoder", ["native", "python"])
def test_generator_reads_stdin_pipe(monkeypatch, decoder):
    counts = synth.synthetic_counts(5, seed=1)
    rows = [utils.format_tensor_row("chr1", 100 + i, "A" * 33, counts[i]) for i in range(5)]
    monkeypatch.setattr(sys, "stdin", io.StringIO("\n".join(rows) + "\n"))
    got = list(utils.tensor_generator_from("PIPE", 2, decoder=decoder))
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


def test_binary_transport_equals_the_reference_text_decode(tmp_path, capsys):
    # the same 11 golden rows through the binary framing: identical batches, infos and progress lines as the text path
    # (whose expectations were produced by the reference's own generator)
    with np.load(os.path.join(GOLDEN, "decode_expected.npz")) as z:
        exp = {k: z[k] for k in z.files}
    with gzip.open(os.path.join(GOLDEN, "decode_rows.txt.gz"), "rt") as f:
        rows = f.read().splitlines()
    p = tmp_path / "t.bin"
    with open(p, "wb") as fo:
        utils.text_to_binary(rows, fo)
    assert os.path.getsize(p) == 8 + 11 * 2192
    got = list(utils.binary_tensor_generator_from(str(p), 4))
    assert len(got) == 3
    for bi, (X, infos) in enumerate(got):
        assert X.dtype == np.int16
        np.testing.assert_array_equal(X.astype(np.float32), exp["X%d" % bi])
        assert [" ".join(i) for i in infos] == list(exp["info%d" % bi])
    assert capsys.readouterr().err == str(exp["stderr"])
    f32 = list(utils.binary_tensor_generator_from(str(p), 100, as_float32=True))
    assert len(f32) == 1 and f32[0][0].dtype == np.float32 and len(f32[0][1]) == 10


def test_binary_transport_errors_and_edges(tmp_path):
    counts = synth.synthetic_counts(3, seed=4)
    p = tmp_path / "x.bin"
    with open(p, "wb") as fo:
        utils.write_binary_tensors(fo, [("chr1", 5 + i, "A" * 16 + "*" + "A" * 16, counts[i]) for i in range(3)])
    assert list(utils.binary_tensor_generator_from(str(p), 2)) == []          # every centre base filtered
    with open(p, "wb") as fo:
        fo.write(utils.BINARY_MAGIC)
    assert list(utils.binary_tensor_generator_from(str(p), 2)) == []          # header only
    with open(p, "wb") as fo:
        utils.write_binary_tensors(fo, [("chr1", 5, "A" * 33, counts[0])])
        fo.write(b"\x00" * 100)
    import pytest
    with pytest.raises(ValueError, match="truncated"):
        list(utils.binary_tensor_generator_from(str(p), 2))
    with open(p, "wb") as fo:
        fo.write(b"chr1 5 AAAA 1 2 3\n")
    with pytest.raises(ValueError, match="not a clair_b200"):
        list(utils.binary_tensor_generator_from(str(p), 2))
    with pytest.raises(ValueError):
        utils.write_binary_tensors(open(p, "wb"), [("c" * 40, 1, "A" * 33, counts[0])])
    pinned = []
    with open(p, "wb") as fo:
        utils.write_binary_tensors(fo, [("chrX", 7 + i, "ACGT" * 8 + "N", counts[i]) for i in range(3)])
    def alloc(shape, dtype):
        pinned.append(np.empty(shape, dtype))
        return pinned[-1]
    got = list(utils.binary_tensor_generator_from(str(p), 2, alloc=alloc))
    assert len(got) == 2 and np.shares_memory(got[0][0], pinned[0])
    np.testing.assert_array_equal(np.concatenate([x for x, _ in got]).astype(np.float32), synth.synthetic_tensors(3, seed=4))


def test_native_decoder_equals_python_decoder_on_awkward_rows(tmp_path):
    # tabs, runs of blanks, CRLF, negative and float-formatted values, a last line without newline, int16 output
    rng = np.random.default_rng(5)
    rows = []
    for i in range(9):
        vals = rng.integers(-40, 300, 1056)
        toks = ["%d" % v for v in vals]
        if i == 2:
            toks[7] = "12.0"
            toks[9] = "1e1"
            toks[11] = "+5"
        sep = "\t" if i % 3 == 0 else ("  " if i % 3 == 1 else " ")
        seq = "ACGTN" * 6 + "ACG"
        if i == 4:
            seq = seq[:16] + "-" + seq[17:]                        # dropped by the IUPAC filter
        rows.append(sep.join(["ctg_%d" % i, str(1000 + i), seq] + toks) + ("\r" if i == 5 else ""))
    p = tmp_path / "rows.gz"
    with gzip.open(p, "wt", newline="") as f:
        f.write("\n".join(rows))                                   # no trailing newline
    a = list(utils.tensor_generator_from(str(p), 4, decoder="native"))
    b = list(utils.tensor_generator_from(str(p), 4, decoder="python"))
    assert len(a) == len(b) == 3
    for (Xa, ia), (Xb, ib) in zip(a, b):
        np.testing.assert_array_equal(Xa, Xb)
        assert ia == ib and Xa.dtype == np.float32
    rows_i = [r for k, r in enumerate(rows) if k != 2]              # int16 transport: integers only
    with gzip.open(p, "wt", newline="") as f:
        f.write("\n".join(rows_i) + "\n")
    c = list(utils.tensor_generator_from(str(p), 100, decoder="native", dtype=np.int16))
    d = list(utils.tensor_generator_from(str(p), 100, decoder="python"))
    assert c[0][0].dtype == np.int16 and c[0][1] == d[0][1]
    np.testing.assert_array_equal(c[0][0].astype(np.float32), d[0][0])


def test_native_decoder_rejects_malformed_rows(tmp_path):
    counts = synth.synthetic_counts(2, seed=3)
    good = utils.format_tensor_row("c", 1, "A" * 33, counts[0])
    for bad in (good.rsplit(" ", 1)[0], good + " 7", "c 1 " + "A" * 10 + good[good.index("A" * 33) + 33:], "",
                good.replace(" 1 ", " x1y ", 1) if False else "c 1"):
        p = tmp_path / "bad.gz"
        with gzip.open(p, "wt") as f:
            f.write(good + "\n" + bad + "\n")
        with pytest.raises(ValueError):
            list(utils.tensor_generator_from(str(p), 4, decoder="native"))
    with gzip.open(p, "wt") as f:
        f.write(good.replace(" %d " % counts[0].reshape(-1)[5], " 40000 ", 1) + "\n")
    with pytest.raises(ValueError, match="int16"):
        list(utils.tensor_generator_from(str(p), 4, decoder="native", dtype=np.int16))
    with pytest.raises(ValueError):
        list(utils.tensor_generator_from(str(p), 4, decoder="python", dtype=np.int16))


def test_native_decoder_randomised_rows_match_python(tmp_path):
    # randomised separators / signs / widths / IUPAC centres, several batches, both decoders must agree exactly
    rng = np.random.default_rng(11)
    seps = [" ", "\t", "  ", " \t "]
    rows = []
    for i in range(57):
        vals = rng.integers(-999, 30000, 1056) if i % 7 == 0 else rng.integers(0, 60, 1056)
        toks = ["%d" % v for v in vals]
        for k in rng.integers(0, 1056, 3):
            toks[k] = "+" + toks[k] if not toks[k].startswith("-") else toks[k]
        centre = "ACGTURYSWKMBDHVNacgtn*-."[int(rng.integers(0, 24))]
        seq = "".join(rng.choice(list("ACGT"), 16)) + centre + "".join(rng.choice(list("ACGT"), 16))
        sep = seps[int(rng.integers(0, len(seps)))]
        lead = " " if i % 5 == 0 else ""
        rows.append(lead + sep.join(["chr%d" % (i % 4), str(10 ** (i % 9) + i), seq] + toks) + (" " if i % 6 == 0 else ""))
    p = tmp_path / "r.gz"
    with gzip.open(p, "wt") as f:
        f.write("\n".join(rows) + "\n")
    a = list(utils.tensor_generator_from(str(p), 10, decoder="native"))
    b = list(utils.tensor_generator_from(str(p), 10, decoder="python"))
    assert len(a) == len(b) > 0
    for (Xa, ia), (Xb, ib) in zip(a, b):
        np.testing.assert_array_equal(Xa, Xb)
        assert ia == ib


def test_native_decoder_refuses_more_rows_than_the_batch_or_the_buffer_holds():
    # round-1 advisor finding: 50 lines with batch_size=2 used to be written into a 2-row buffer
    from clair_b200 import synth
    c = synth.synthetic_counts(50, seed=1)
    lines = [utils.format_tensor_row("c", i, "A" * 33, c[i]) + "\n" for i in range(50)]
    with pytest.raises(ValueError):
        utils.native_rows_to_batch(lines, 2)
    with pytest.raises(ValueError):
        utils.native_rows_to_batch(lines, 50, out=np.empty((10, 1056), np.float32))
    with pytest.raises(ValueError):
        utils.native_rows_to_batch(lines, 50, out=np.empty((50, 1056), np.float64))
    X, infos = utils.native_rows_to_batch(lines, 50, out=np.empty((64, 1056), np.float32))
    assert X.shape == (50, 33, 8, 4) and len(infos) == 50
