"""Binary dataset ("bin") reader without python-blosc: Blosc1 / LZ4 frame decoder on frames assembled
from the format specification (the real library is absent: codec parity unpinned), and the walk over frames against the
reference's own decompress_array (tests/golden/bins_walk.json, oracle/gen_golden_bins.py)."""
import json
import os
import pickle
import struct
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clair_b200 import bins                                      # noqa: E402
from oracle import gen_golden_bins as G                          # noqa: E402


def test_lz4_known_blocks():
    """Hand-assembled LZ4 blocks inside stored-size Blosc frames: literals only, a short match, an overlapping (run-length)
    match, extended literal and match lengths."""
    def frame(block, nbytes):
        # one block, not split (flag 0x10), codec LZ4 (1 << 5)
        body = struct.pack("<i", len(block)) + block
        return struct.pack("<BBBBIII", 2, 1, 0x10 | (1 << 5), 1, nbytes, nbytes, 16 + 4 + len(body)) + struct.pack("<i", 20) + body

    lit = bytes(range(200))
    block = bytes([0xF0, 200 - 15]) + lit                                       # 200 literals, extended length
    assert bins.blosc_decompress(frame(block, 200)) == lit
    # "abcdabcdabcd...": 4 literals then a match of 196 at offset 4 (overlapping copy), then 5 trailing literals
    want = b"abcd" * 50 + b"vwxyz"
    block = bytes([0x4F]) + b"abcd" + struct.pack("<H", 4) + bytes([196 - 4 - 15]) + bytes([0x50]) + b"vwxyz"
    assert bins.blosc_decompress(frame(block, len(want))) == want
    # a run of one byte: 1 literal, match of 299 at offset 1 (two extension bytes), 5 literals
    want = b"z" * 300 + b"12345"
    block = bytes([0x1F]) + b"z" + struct.pack("<H", 1) + bytes([255, 299 - 4 - 15 - 255]) + bytes([0x50]) + b"12345"
    assert bins.blosc_decompress(frame(block, len(want))) == want


def test_corrupt_and_unsupported_frames_are_errors():
    good = bins.blosc_compress(b"0123456789" * 100)
    assert bins.blosc_decompress(good) == b"0123456789" * 100
    with pytest.raises(ValueError):
        bins.blosc_decompress(good[:10])                                          # shorter than a header
    with pytest.raises(ValueError):
        bins.blosc_decompress(good[:-7])                                          # truncated
    zstd = bytearray(good)
    zstd[2] = (zstd[2] & 0x1F) | (4 << 5)
    with pytest.raises(ValueError):
        bins.blosc_decompress(bytes(zstd))                                        # another codec
    bad = bytearray(good)
    bad[16 + 4 + 4 + 5] ^= 0xFF                                                  # damage inside the LZ4 stream
    try:
        assert bins.blosc_decompress(bytes(bad)) != b"0123456789" * 100
    except ValueError:
        pass
    offset0 = struct.pack("<BBBBIII", 2, 1, 0x10 | (1 << 5), 1, 200, 200, 16 + 4 + 4 + 9) + struct.pack("<i", 20) + \
        struct.pack("<i", 9) + bytes([0x1F]) + b"z" + struct.pack("<H", 0) + bytes([180]) + bytes([0x00])
    with pytest.raises(ValueError):
        bins.blosc_decompress(offset0)                                            # match offset 0


@pytest.mark.parametrize("split", [True, False])
@pytest.mark.parametrize("typesize,n", [(4, 100000), (4, 262144 + 999), (8, 70001), (1, 5000), (2, 127), (16, 40000), (80, 9000)])
def test_frames_round_trip(typesize, n, split):
    rng = np.random.default_rng(n + typesize)
    data = (rng.integers(0, 4, size=n).astype(np.uint8) * 17).tobytes()           # compressible
    frame = bins.blosc_compress(data, typesize=typesize, split=split)
    assert len(frame) < len(data) or n < 128
    assert bins.blosc_decompress(frame) == data
    noise = rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()               # incompressible: streams are stored
    assert bins.blosc_decompress(bins.blosc_compress(noise, typesize=typesize, split=split)) == noise
    assert bins.blosc_decompress(bins.blosc_compress(data, typesize=typesize, store=True)) == data


def test_shuffled_frame_is_unshuffled():
    """A byte-shuffled, stored-stream frame built by hand (the reference writes NOSHUFFLE; the flag is still honoured)."""
    items = np.arange(300, dtype=np.uint32)
    raw = items.tobytes()
    shuffled = items.view(np.uint8).reshape(-1, 4).T.tobytes()
    body = struct.pack("<i", len(shuffled)) + shuffled
    frame = struct.pack("<BBBBIII", 2, 1, 0x1 | 0x10 | (1 << 5), 4, len(raw), len(raw), 16 + 4 + len(body)) + \
        struct.pack("<i", 20) + body
    assert bins.blosc_decompress(frame) == raw


def test_pack_unpack_arrays():
    from clair_b200 import synth
    x = synth.synthetic_counts(60, seed=5).astype(np.float32)
    y = np.eye(90, dtype=np.float32)[np.arange(60) % 90]
    pos = np.array(["chr1:%d" % (1000 + i) for i in range(60)])
    for a in (x, y, pos):
        back = bins.unpack_array(bins.pack_array(a))
        assert back.dtype == a.dtype and np.array_equal(back, a)


def test_walk_matches_reference_decompress_array():
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "bins_walk.json")))
    assert len(golden) >= 18
    for g in golden:
        frames = G.frames_of(g["sizes"])
        assert G.walk(bins.decompress_array, frames, g["batch"], g.get("read_index_list")) == g["steps"]


def test_bin_file_round_trip(tmp_path):
    from clair_b200 import synth
    n = 1234
    x = synth.synthetic_counts(n, seed=9).astype(np.float32)
    y = np.eye(90, dtype=np.float32)[np.arange(n) % 90]
    pos = np.array(["chr1:%d" % i for i in range(n)])
    path = str(tmp_path / "tensors.bin")
    bins.write_bin(path, x, y, pos)
    with open(path, "rb") as fh:                                                  # Tensor2Bin.py:28-33 layout
        assert pickle.load(fh) == n and len(pickle.load(fh)) == 3
    info = bins.dataset_info_from(binary_file_path=path)
    assert info.dataset_size == n and not info.is_separated_train_and_validation_binary
    assert bins.no_of_blosc_blocks_from(info, int(n * 0.9)) == 3
    got = list(bins.prediction_batches_from(info, 500))
    assert [len(xb) for xb, _ in got] == [500, 500, 234]
    assert np.array_equal(np.concatenate([xb for xb, _ in got]), x) and np.array_equal(np.concatenate([yb for _, yb in got]), y)
    two = bins.dataset_info_from(train_binary_file_path=path, validation_binary_file_path=path)
    assert two.dataset_size == 2 * n and two.no_of_training_examples_from_train_binary == n
    assert bins.no_of_blosc_blocks_from(two, n) == 6


@pytest.mark.gpu
def test_bin_batches_through_predict(tmp_path):
    """evaluate.py:64-82: batches from a bin go straight into Clair.predict."""
    from clair_b200 import synth, weights as W
    from clair_b200.model import Clair
    n = 700
    x = synth.synthetic_tensors(n, seed=11)
    path = str(tmp_path / "tensors.bin")
    bins.write_bin(path, x, np.zeros((n, 90), np.float32), np.array(["c:%d" % i for i in range(n)]))
    m = Clair(max_sites=2048, batch_sites=1000)
    m.set_weights(W.random_weights(seed=1234))
    want = m.predict_packed(x)
    got = np.concatenate([np.concatenate(m.predict(xb), axis=1) for xb, _ in bins.prediction_batches_from(bins.dataset_info_from(path), 500)])
    m.close()
    assert np.array_equal(got, want)


def test_evaluate_model_prints_the_reference_report(tmp_path):
    """clair_b200.evaluate.evaluate_model against what the reference's own evaluate_model printed for the same seeded bin and
    stand-in model (tests/golden/evaluate_report.txt, oracle/gen_golden_evaluate.py)."""
    from clair_b200 import evaluate
    from oracle import gen_golden_evaluate as GE
    info = GE.dataset_info(str(tmp_path))
    assert info.dataset_size == GE.N_SITES and len(info.x_array_compressed) == 4
    want = open(os.path.join(ROOT, "tests", "golden", "evaluate_report.txt")).read()
    assert GE.report_of(evaluate.evaluate_model, info) == want
    model = GE.FakeModel()
    evaluate.evaluate_model(model, info)
    assert model.calls == [1000, 730]                     # predict sees the batches the reference's walk produces


def test_decoder_survives_damaged_frames():
    """Every byte of the container is attacker-controlled as far as the decoder is concerned: random damage must end in an
    error or in bytes of the announced size, never in a read or write outside the buffers (the library would take the test
    process down with it)."""
    rng = np.random.default_rng(123)
    data = (rng.integers(0, 6, size=70000).astype(np.uint8) * 9).tobytes()
    frames = [bins.blosc_compress(data, typesize=4, split=True), bins.blosc_compress(data, typesize=4, split=False),
              bins.blosc_compress(data[:5000], typesize=1)]
    outcomes = {"error": 0, "bytes": 0}
    for frame in frames:
        for _ in range(400):
            damaged = bytearray(frame)
            for _ in range(int(rng.integers(1, 6))):
                damaged[int(rng.integers(0, len(damaged)))] = int(rng.integers(0, 256))
            if rng.random() < 0.2:
                damaged = damaged[:int(rng.integers(0, len(damaged)))]
            try:
                out = bins.blosc_decompress(bytes(damaged))
                assert isinstance(out, bytes)
                outcomes["bytes"] += 1
            except (ValueError, MemoryError):
                outcomes["error"] += 1
    assert outcomes["error"] > 100 and outcomes["bytes"] > 100
