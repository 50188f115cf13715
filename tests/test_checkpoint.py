"""TF-V2 checkpoint bundle reader: table format, protobuf entries, checksums, error paths.
Parity unpinned against a real TensorFlow bundle (none ships with the reference); the table-format corner cases are
checked on hand-assembled bytes, the rest through the writer that follows the same specification."""
import struct

import numpy as np
import pytest

from clair_b200 import checkpoint as C, weights as W


def test_crc32c_known_answers():
    assert C.crc32c(b"123456789") == 0xE3069283                     # RFC 3720 B.4 check value
    assert C.crc32c(b"") == 0
    assert C.crc32c(bytes(32)) == 0x8A9136AA                        # RFC 3720 B.4: 32 bytes of zeros
    assert C.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43               # 32 bytes of ones
    assert C.crc32c(bytes(range(32))) == 0x46DD794E                 # incrementing
    assert C.unmask_crc(C.mask_crc(0xDEADBEEF)) == 0xDEADBEEF


def test_crc32c_lane_path_equals_scalar_path():
    rng = np.random.default_rng(0)
    for n in (4096, 5000, 70001, 300007):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        ref = 0xFFFFFFFF
        t = C._crc_table()
        for b in data:
            ref = int(t[(ref ^ b) & 0xFF]) ^ (ref >> 8)
        assert C.crc32c(data) == ref ^ 0xFFFFFFFF, n


def test_hand_assembled_block_with_prefix_compression():
    # entries: "L4/bias"->b"x", "L4/kernel"->b"yz" (shares "L4/"), restart array [0], count 1
    e1 = bytes([0, 7, 1]) + b"L4/bias" + b"x"
    e2 = bytes([3, 6, 2]) + b"kernel" + b"yz"
    block = e1 + e2 + struct.pack("<II", 0, 1)
    assert list(C._block_entries(block)) == [(b"L4/bias", b"x"), (b"L4/kernel", b"yz")]
    with pytest.raises(ValueError):
        list(C._block_entries(bytes([9, 1, 0]) + b"a" + struct.pack("<II", 0, 1)))      # shared > previous key


def test_bundle_entry_proto_fields():
    # dtype=1, shape {dim{size:7680} dim{size:192}}, offset 1234, size 5898240, crc fixed32
    dims = b"".join(bytes([0x12]) + C._put_varint(len(d)) + d for d in (bytes([0x08]) + C._put_varint(s) for s in (7680, 192)))
    buf = bytes([0x08, 1, 0x12]) + C._put_varint(len(dims)) + dims + bytes([0x20]) + C._put_varint(1234) + \
        bytes([0x28]) + C._put_varint(5898240) + bytes([0x35]) + struct.pack("<I", 0xAABBCCDD)
    e = C.parse_bundle_entry(buf)
    assert e == {"dtype": 1, "shape": (7680, 192), "shard_id": 0, "offset": 1234, "size": 5898240,
                 "crc32c": 0xAABBCCDD, "sliced": False}


@pytest.mark.parametrize("block_size,restart", [(4096, 16), (256, 3), (64, 1)])
def test_round_trip_full_model(tmp_path, block_size, restart):
    w = W.random_weights(seed=3)
    extra = dict(w)
    extra["global_step"] = np.array(7, dtype=np.int64)              # non-float entries and optimiser slots are ignored
    extra["L4/kernel/Adam"] = np.zeros((7680, 192), np.float32)
    extra["beta1_power"] = np.array(0.9, dtype=np.float32)
    prefix = str(tmp_path / "model")
    C.write_bundle(prefix, extra, block_size=block_size, restart_interval=restart)
    header, entries = C.read_index(prefix + ".index")
    assert header["num_shards"] == 1 and len(entries) == len(extra)
    assert list(entries) == sorted(entries, key=lambda s: s.encode())   # table keys are sorted
    got = C.load_checkpoint(prefix)
    assert set(got) == set(w)
    for k in w:
        np.testing.assert_array_equal(got[k], w[k])
    assert C.is_checkpoint_prefix(prefix) and C.is_checkpoint_prefix(prefix + ".index")
    everything = C.read_bundle(prefix)
    assert everything["global_step"] == 7 and everything["global_step"].dtype == np.int64


def test_corruption_and_missing_variables_are_errors(tmp_path):
    w = W.random_weights(seed=4)
    prefix = str(tmp_path / "m")
    C.write_bundle(prefix, w)
    data_path = prefix + ".data-00000-of-00001"
    raw = bytearray(open(data_path, "rb").read())
    raw[1000] ^= 0x40
    open(data_path, "wb").write(bytes(raw))
    with pytest.raises(ValueError, match="checksum"):
        C.load_checkpoint(prefix)
    C.load_checkpoint(prefix, verify=False)                          # explicit opt-out still reads
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[10] ^= 0x01
    open(prefix + ".index", "wb").write(bytes(idx))
    with pytest.raises(ValueError):
        C.read_index(prefix + ".index")
    idx[-1] ^= 0xFF
    open(prefix + ".index", "wb").write(bytes(idx))
    with pytest.raises(ValueError, match="magic"):
        C.read_index(prefix + ".index")
    del w["L4/bias"]
    prefix2 = str(tmp_path / "m2")
    C.write_bundle(prefix2, w)
    with pytest.raises(ValueError, match="L4/bias"):
        C.load_checkpoint(prefix2)


@pytest.mark.gpu
def test_restore_parameters_from_a_bundle(tmp_path, weights1234):
    from clair_b200.model import Clair
    from clair_b200 import synth
    prefix = str(tmp_path / "model")
    C.write_bundle(prefix, weights1234)
    a, b = Clair(max_sites=512), Clair(max_sites=512)
    a.restore_parameters(prefix)                                     # like m.restore_parameters(chkpnt_fn), call_var.py:215
    b.set_weights(weights1234)
    X = synth.synthetic_tensors(200, seed=3)
    np.testing.assert_array_equal(a.predict_packed(X), b.predict_packed(X))
    a.close()
    b.close()


# ---- LSTM variable discovery (the released models were saved through the cuDNN branch, clair/model.py:281-296) -------------
def _renamed(w, rename):
    return {rename(k): v for k, v in w.items()}


@pytest.mark.parametrize("scheme", ["cpu_graph", "cudnn_layer_scope", "other_cell_name", "opaque_buffer"])
def test_lstm_variables_are_discovered_under_every_naming_scheme(tmp_path, scheme):
    w = W.random_weights(seed=11)
    lstm_keys = [k for k in w if k.startswith("LSTM")]
    if scheme == "cpu_graph":
        saved = dict(w)
    elif scheme == "cudnn_layer_scope":            # canonical tensors written by the CudnnLSTM layer's saveable, under its scope
        saved = _renamed(w, lambda k: k.replace("/stack_bidirectional_rnn/", "/cudnn_lstm/stack_bidirectional_rnn/") if k in lstm_keys else k)
    elif scheme == "other_cell_name":
        saved = _renamed(w, lambda k: k.replace("cudnn_compatible_lstm_cell", "lstm_cell") if k in lstm_keys else k)
    else:                                          # only the raw cuDNN parameter buffer of each layer
        saved = {k: v for k, v in w.items() if k not in lstm_keys}
        for layer, fin in ((1, 32), (2, 256)):
            parts_w, parts_b = [], []
            for d in ("fw", "bw"):
                k, b = w[W.lstm_name(layer, d, "kernel")], w[W.lstm_name(layer, d, "bias")]
                gates = np.split(k, 4, axis=1)                     # TF columns i, c, f, o
                cudnn = [gates[0], gates[2], gates[1], gates[3]]   # cuDNN order i, f, c, o
                parts_w += [g[:fin].T.reshape(-1) for g in cudnn] + [g[fin:].T.reshape(-1) for g in cudnn]
                bg = np.split(b, 4)
                bc = [bg[0], bg[2], bg[1], bg[3]]
                parts_b += [0.25 * x for x in bc] + [0.75 * x for x in bc]      # b_W + b_R = the TF bias
            saved["LSTM%d/cudnn_lstm/opaque_kernel" % layer] = np.concatenate(parts_w + parts_b).astype(np.float32)
    # what every training checkpoint also holds: optimiser slots next to the variables, step counters
    for k in list(saved):
        if k.endswith("/kernel") and saved[k].ndim == 2:
            saved[k + "/Adam"] = np.zeros_like(saved[k])
            saved[k + "/Adam_1"] = np.zeros_like(saved[k])
    saved["global_step"] = np.array(3, dtype=np.int64)
    prefix = str(tmp_path / scheme)
    C.write_bundle(prefix, saved)
    got = C.load_checkpoint(prefix)
    assert set(got) == set(w)
    for k in w:
        if scheme == "opaque_buffer" and k.endswith("/bias") and k.startswith("LSTM"):
            np.testing.assert_allclose(got[k], w[k], rtol=1e-6, atol=1e-7)      # 0.25 b + 0.75 b in float32
        else:
            np.testing.assert_array_equal(got[k], w[k], err_msg=k)


def test_ambiguous_or_missing_lstm_variables_are_reported(tmp_path):
    w = W.random_weights(seed=12)
    twice = dict(w)
    k = W.lstm_name(1, "fw", "kernel")
    del twice[k]
    twice["LSTM1/a/fw/cell/kernel"] = w[k]
    twice["LSTM1/b/fw/cell/kernel"] = w[k]
    C.write_bundle(str(tmp_path / "twice"), twice)
    with pytest.raises(ValueError, match="candidates"):
        C.load_checkpoint(str(tmp_path / "twice"))
    gone = {n: v for n, v in w.items() if n != W.lstm_name(2, "bw", "bias")}
    C.write_bundle(str(tmp_path / "gone"), gone)
    with pytest.raises(ValueError, match="cannot identify LSTM2"):
        C.load_checkpoint(str(tmp_path / "gone"))
    with pytest.raises(ValueError):
        C.cudnn_opaque_to_canonical(np.zeros(17, np.float32), 32, 128)
