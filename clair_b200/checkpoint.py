"""TensorFlow "V2" checkpoint bundles -> weight blob, without TensorFlow.

The reference restores its models with ``tf.train.Saver.restore(session, prefix)`` (clair/model.py:712, 1016-1020);
the published ONT / PacBio / Illumina models (README.md:359-365) are TF-1.13 V2 bundles:
    <prefix>.index                    an SSTable (LevelDB table format) mapping variable name -> BundleEntryProto;
                                      the empty key holds the BundleHeaderProto
    <prefix>.data-00000-of-0000N      raw little-endian tensor bytes, addressed by (shard_id, offset, size)
TensorFlow is not installable here, so this follows the published formats:
    table format  tensorflow/core/lib/io/format.{h,cc}, block.cc, table_builder.cc (= LevelDB doc/table_format.md):
                  footer 48 B = metaindex handle | index handle | padding | magic 0xdb4775248b80fb57; a block is
                  prefix-compressed entries + restart array + count, followed by 1 type byte and a masked CRC32C
    protos        tensorflow/core/protobuf/tensor_bundle.proto (BundleHeaderProto, BundleEntryProto),
                  tensor_shape.proto, types.proto (DT_FLOAT = 1)
    checksums     tensorflow/core/lib/hash/crc32c.h: CRC32C (Castagnoli), mask = rotr15(crc) + 0xa282ead8
PARITY UNPINNED against a real bundle: the reference ships no checkpoint and TensorFlow cannot run here; the reader is
tested against bundles produced by `write_bundle` below (same specification, so a shared misreading would go
unnoticed) and against hand-assembled byte strings for the table-format corner cases (prefix compression, several
data blocks, restart points).  Variable names expected in the bundle are those of clair_b200.weights.weight_shapes();
optimiser slots (".../Adam", "beta1_power", ...) and non-float entries are ignored.
"""
import os
import struct

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
FOOTER_LEN = 48
BLOCK_TRAILER_LEN = 5
DT_FLOAT = 1
_NP_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64}       # types.proto

# ---- CRC32C ---------------------------------------------------------------------------------------------------------
_CRC_TABLE = None


def _crc_table():
    global _CRC_TABLE
    if _CRC_TABLE is None:
        poly = 0x82F63B78
        t = np.zeros(256, dtype=np.uint32)
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ poly if c & 1 else c >> 1
            t[i] = c
        _CRC_TABLE = t
    return _CRC_TABLE


def crc32c(data):
    """CRC32C (Castagnoli) of a bytes-like object.  Large buffers are folded in parallel lanes with numpy and the
    lane CRCs combined through the zero-extension operator (CRC is linear over GF(2))."""
    buf = np.frombuffer(bytes(data) if not isinstance(data, (bytes, bytearray, memoryview)) else data, dtype=np.uint8)
    t = _crc_table()
    n = len(buf)
    if n < 4096:
        c = 0xFFFFFFFF
        for b in buf.tolist():
            c = int(t[(c ^ b) & 0xFF]) ^ (c >> 8)
        return c ^ 0xFFFFFFFF
    lanes = max(16, min(2048, n // 512))
    seg = n // lanes
    body = buf[:seg * lanes].reshape(lanes, seg)
    c = np.zeros(lanes, dtype=np.uint32)
    c[0] = 0xFFFFFFFF                                  # only the first lane carries the initial value
    for k in range(seg):
        c = t[(c ^ body[:, k]) & 0xFF] ^ (c >> 8)
    # combine: crc(A || B) = shift(crc_state(A), len(B)) xor crc_state0(B), shift = feeding len(B) zero bytes
    total = int(c[0])
    for i in range(1, lanes):
        total = _crc_shift(total, seg) ^ int(c[i])
    for b in buf[seg * lanes:].tolist():
        total = int(t[(total ^ b) & 0xFF]) ^ (total >> 8)
    return total ^ 0xFFFFFFFF


def _gf2_matrix_times(mat, vec):
    s = 0
    i = 0
    while vec:
        if vec & 1:
            s ^= mat[i]
        vec >>= 1
        i += 1
    return s


def _gf2_matrix_square(mat):
    return [_gf2_matrix_times(mat, mat[i]) for i in range(32)]


_SHIFT_CACHE = {}


def _crc_shift(crc, nbytes):
    """State after feeding `nbytes` zero bytes (zlib's crc32_combine construction, reflected CRC32C polynomial)."""
    if nbytes == 0:
        return crc
    ops = _SHIFT_CACHE.get(nbytes)
    if ops is None:
        odd = [0x82F63B78] + [1 << i for i in range(31)]      # operator for one zero bit
        even = _gf2_matrix_square(odd)                          # two bits
        odd = _gf2_matrix_square(even)                          # four bits
        ops = []
        n = nbytes
        # first square puts the operator for one zero byte (8 bits) in `even`
        while True:
            even = _gf2_matrix_square(odd)
            if n & 1:
                ops.append(even)
            n >>= 1
            if not n:
                break
            odd = _gf2_matrix_square(even)
            if n & 1:
                ops.append(odd)
            n >>= 1
            if not n:
                break
        _SHIFT_CACHE[nbytes] = ops
    for m in ops:
        crc = _gf2_matrix_times(m, crc)
    return crc


def mask_crc(crc):
    return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def unmask_crc(masked):
    rot = (masked - 0xA282EAD8) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ---- varints / protobuf wire format -----------------------------------------------------------------------------------
def _get_varint(buf, pos):
    result = shift = 0
    while True:
        if pos >= len(buf):
            raise ValueError("truncated varint")
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 70:
            raise ValueError("varint too long")


def _put_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_proto(buf):
    """Minimal protobuf reader: {field_number: [values]} with varints as int, length-delimited as bytes,
    fixed32/64 as int."""
    fields = {}
    pos = 0
    while pos < len(buf):
        key, pos = _get_varint(buf, pos)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + ln])
            if len(v) != ln:
                raise ValueError("truncated protobuf field")
            pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        fields.setdefault(fn, []).append(v)
    return fields


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def parse_bundle_entry(buf):
    """BundleEntryProto -> dict(dtype, shape, shard_id, offset, size, crc32c, sliced)."""
    f = _parse_proto(buf)
    shape = []
    for sp in f.get(2, []):                                   # TensorShapeProto
        sf = _parse_proto(sp)
        for dim in sf.get(2, []):                             # repeated Dim
            df = _parse_proto(dim)
            shape.append(_signed64(df.get(1, [0])[0]))
    return {"dtype": f.get(1, [0])[0], "shape": tuple(shape), "shard_id": f.get(3, [0])[0],
            "offset": f.get(4, [0])[0], "size": f.get(5, [0])[0], "crc32c": f.get(6, [None])[0],
            "sliced": 7 in f}


def parse_bundle_header(buf):
    f = _parse_proto(buf)
    return {"num_shards": f.get(1, [0])[0], "endianness": f.get(2, [0])[0]}     # 0 = LITTLE


# ---- table reader -------------------------------------------------------------------------------------------------------
def _read_block(data, offset, size, verify):
    end = offset + size
    if end + BLOCK_TRAILER_LEN > len(data):
        raise ValueError("block handle points outside the file")
    block = data[offset:end]
    ctype = data[end]
    if verify:
        stored = struct.unpack_from("<I", data, end + 1)[0]
        if unmask_crc(stored) != crc32c(data[offset:end + 1]):
            raise ValueError("index block checksum mismatch at offset %d" % offset)
    if ctype != 0:
        raise ValueError("compressed table blocks (type %d) are not supported" % ctype)
    return block


def _block_entries(block):
    """Yield (key, value) of one table block (prefix-compressed keys; the restart array is only an index)."""
    if len(block) < 4:
        raise ValueError("table block too short")
    num_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * num_restarts
    if limit < 0:
        raise ValueError("bad restart array")
    pos = 0
    key = b""
    while pos < limit:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        if shared > len(key) or pos + non_shared + vlen > limit:
            raise ValueError("corrupt table entry")
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        value = bytes(block[pos:pos + vlen])
        pos += vlen
        yield key, value


def read_index(path, verify=True):
    """<prefix>.index -> (header dict, {name: entry dict}) in key order."""
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < FOOTER_LEN:
        raise ValueError("%s: too short for a table footer" % path)
    footer = data[-FOOTER_LEN:]
    if struct.unpack_from("<Q", footer, 40)[0] != TABLE_MAGIC:
        raise ValueError("%s: not a TensorFlow table (bad magic)" % path)
    pos = 0
    _, pos = _get_varint(footer, pos)           # metaindex handle
    _, pos = _get_varint(footer, pos)
    idx_off, pos = _get_varint(footer, pos)
    idx_size, pos = _get_varint(footer, pos)
    header, entries = None, {}
    for _, handle in _block_entries(_read_block(data, idx_off, idx_size, verify)):
        off, p = _get_varint(handle, 0)
        size, p = _get_varint(handle, p)
        for key, value in _block_entries(_read_block(data, off, size, verify)):
            if key == b"":
                header = parse_bundle_header(value)
            else:
                entries[key.decode("utf-8")] = parse_bundle_entry(value)
    if header is None:
        raise ValueError("%s: no bundle header entry" % path)
    if header["endianness"] != 0:
        raise ValueError("big-endian bundles are not supported")
    return header, entries


def _shard_path(prefix, shard, num_shards):
    return "%s.data-%05d-of-%05d" % (prefix, shard, num_shards)


def read_bundle(prefix, names=None, verify=True):
    """{name: ndarray} of the (float) tensors of a V2 bundle.  `names`: restrict to these variables."""
    prefix = str(prefix)
    if prefix.endswith(".index"):
        prefix = prefix[:-len(".index")]
    header, entries = read_index(prefix + ".index", verify)
    shards = {}
    out = {}
    for name, e in entries.items():
        if names is not None and name not in names:
            continue
        if e["sliced"] or e["dtype"] not in _NP_DTYPES:
            if names is not None:
                raise ValueError("variable %s is stored sliced or in an unsupported dtype" % name)
            continue
        sid = e["shard_id"]
        if sid not in shards:
            shards[sid] = np.memmap(_shard_path(prefix, sid, header["num_shards"]), dtype=np.uint8, mode="r")
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        dt = np.dtype(_NP_DTYPES[e["dtype"]])
        count = int(np.prod(e["shape"])) if e["shape"] else 1
        if len(raw) != e["size"] or count * dt.itemsize != e["size"]:
            raise ValueError("variable %s: size %d does not match shape %s" % (name, e["size"], e["shape"]))
        if verify and e["crc32c"] is not None and unmask_crc(e["crc32c"]) != crc32c(raw):
            raise ValueError("variable %s: data checksum mismatch" % name)
        out[name] = np.frombuffer(bytes(raw), dtype=dt.newbyteorder("<")).reshape(e["shape"]).astype(dt)
    return out


# ---- which entries are the forward path's variables ------------------------------------------------------------------------
# The dense layers carry explicit names (clair/model.py:240, 482-618: "L3/Unit_c", "L4", "L5_k", "Prediction/..."), but the
# names of the LSTM variables depend on which branch of adaptive_LSTM_layer built the graph that was SAVED:
#   * CPU branch (clair/model.py:298-312): "LSTM1/stack_bidirectional_rnn/cell_0/bidirectional_rnn/{fw,bw}/
#     cudnn_compatible_lstm_cell/{kernel,bias}" - what SURVEY.md 8a lists and weights.lstm_name() produces;
#   * cuDNN branch (clair/model.py:281-296, the one the released models were trained with): the CudnnLSTM layer keeps ONE
#     opaque parameter buffer whose saveable writes either the same canonical tensors under the layer's own scope
#     ("LSTM1/cudnn_lstm/stack_bidirectional_rnn/.../kernel") or, saved raw, the buffer itself ("LSTM1/cudnn_lstm/opaque_kernel").
# So the LSTM variables are DISCOVERED: by scope ("LSTM1/" ... "/fw/" ... "/kernel"), shape and - last - the opaque buffer.
_SLOT_SUFFIXES = ("/Adam", "/Adam_1", "/Momentum", "/RMSProp", "/RMSProp_1", "/Adagrad", "/ExponentialMovingAverage")


def _is_slot(name):
    return name.endswith(_SLOT_SUFFIXES)


def discover_lstm_names(entries):
    """{canonical name (weights.lstm_name): name in the checkpoint} for the 8 LSTM variables, or {canonical: ("opaque",
    name, direction index)} where only the cuDNN parameter buffer of the layer was saved.  Raises ValueError naming the
    candidates when a variable cannot be identified uniquely."""
    from . import weights as W
    found = {}
    for layer, fin in ((1, W.F), (2, 2 * W.H)):
        scope = "LSTM%d/" % layer
        inside = [n for n in entries if n.startswith(scope) and not _is_slot(n) and entries[n]["dtype"] == DT_FLOAT]
        for di, d in enumerate(("fw", "bw")):
            for var, shape in (("kernel", (fin + W.H, 4 * W.H)), ("bias", (4 * W.H,))):
                canonical = W.lstm_name(layer, d, var)
                if canonical in entries:
                    found[canonical] = canonical
                    continue
                cands = [n for n in inside if n.endswith("/" + var) and ("/%s/" % d) in n and tuple(entries[n]["shape"]) == shape]
                if len(cands) == 1:
                    found[canonical] = cands[0]
                    continue
                opaque = [n for n in inside if n.endswith("opaque_kernel") and
                          tuple(entries[n]["shape"]) == (cudnn_opaque_size(fin, W.H),)]
                if not cands and len(opaque) == 1:
                    found[canonical] = ("opaque", opaque[0], di)
                    continue
                raise ValueError("cannot identify %s in the checkpoint: %s under %r" % (
                    canonical, ("candidates %s" % cands) if cands else "no variable of shape %s" % (shape,), scope))
    return found


def cudnn_opaque_size(fin, units):
    """Floats in the parameter buffer of a one-layer bidirectional cuDNN LSTM: per direction 4 input matrices [units, fin],
    4 recurrent matrices [units, units] and 8 bias vectors [units]."""
    return 2 * (4 * units * fin + 4 * units * units + 8 * units)


def cudnn_opaque_to_canonical(buf, fin, units):
    """cuDNN LSTM parameter buffer (one bidirectional layer) -> [(kernel [(fin+units), 4 units], bias [4 units]) for fw, bw]
    in the TF LSTMBlockCell / CudnnCompatibleLSTMCell layout.

    Buffer layout (cuDNN canonical order, as TensorFlow 1.13's cudnn_rnn_ops.CudnnOpaqueParamsSaveable unpacks it): all
    weights first - per direction W_i, W_f, W_c, W_o ([units, fin] each, row-major) then R_i, R_f, R_c, R_o ([units, units]) -
    then all biases - per direction b_Wi, b_Wf, b_Wc, b_Wo, b_Ri, b_Rf, b_Rc, b_Ro.  The TF cell wants ONE kernel with rows
    [x ; h] and gate columns i, c, f, o, and ONE bias = b_W + b_R (tf.contrib.cudnn_rnn: _cudnn_to_tf_weights / _biases).
    Unverified against a cuDNN-written buffer (none ships with the reference)."""
    buf = np.asarray(buf, dtype=np.float32).reshape(-1)
    if buf.size != cudnn_opaque_size(fin, units):
        raise ValueError("opaque LSTM buffer holds %d floats, expected %d" % (buf.size, cudnn_opaque_size(fin, units)))
    per_dir_w = 4 * units * fin + 4 * units * units
    out = []
    for di in range(2):
        w = buf[di * per_dir_w:(di + 1) * per_dir_w]
        W_ = w[:4 * units * fin].reshape(4, units, fin)                    # i, f, c, o
        R_ = w[4 * units * fin:].reshape(4, units, units)
        b = buf[2 * per_dir_w + di * 8 * units:2 * per_dir_w + (di + 1) * 8 * units].reshape(2, 4, units)
        order = (0, 2, 1, 3)                                               # TF gate order i, c, f, o
        kernel = np.concatenate([np.concatenate([W_[g].T for g in order], axis=1),
                                 np.concatenate([R_[g].T for g in order], axis=1)], axis=0)
        bias = np.concatenate([b[0, g] + b[1, g] for g in order])
        out.append((np.ascontiguousarray(kernel, dtype=np.float32), np.ascontiguousarray(bias, dtype=np.float32)))
    return out


def load_checkpoint(prefix, verify=True):
    """The forward path's variables of a reference checkpoint as a float32 weight blob keyed by the canonical names
    (what Clair.restore_parameters feeds to clairb_set_weight).  Dense variables are taken by name, LSTM variables are
    discovered (discover_lstm_names)."""
    from . import weights as W
    shapes = W.weight_shapes()
    _, entries = read_index((str(prefix)[:-6] if str(prefix).endswith(".index") else str(prefix)) + ".index", verify)
    lstm = discover_lstm_names(entries)
    missing = [k for k in shapes if k not in entries and k not in lstm]
    if missing:
        raise ValueError("checkpoint %s lacks %d forward-path variables, e.g. %s (has e.g. %s)"
                         % (prefix, len(missing), missing[0], sorted(entries)[:3]))
    source = {k: (lstm[k] if k in lstm else k) for k in shapes}
    wanted = {v if isinstance(v, str) else v[1] for v in source.values()}
    raw = read_bundle(prefix, names=wanted, verify=verify)
    w, opaque_cache = {}, {}
    for k, v in source.items():
        if isinstance(v, str):
            w[k] = np.ascontiguousarray(raw[v], dtype=np.float32)
            continue
        _, name, di = v
        if name not in opaque_cache:
            layer = 1 if name.startswith("LSTM1/") else 2
            opaque_cache[name] = cudnn_opaque_to_canonical(raw[name], W.F if layer == 1 else 2 * W.H, W.H)
        w[k] = opaque_cache[name][di][0 if k.endswith("/kernel") else 1]
    W.check_weights(w)
    return w


def is_checkpoint_prefix(path):
    path = str(path)
    return os.path.exists(path + ".index") or (path.endswith(".index") and os.path.exists(path))


# ---- writer (tests and export; same specification as the reader) ---------------------------------------------------------
def _proto_field(fn, wt, payload):
    return _put_varint((fn << 3) | wt) + payload


def _entry_proto(dtype, shape, shard_id, offset, size, crc):
    dims = b"".join(_proto_field(2, 2, _put_varint(len(d)) + d)
                    for d in (_proto_field(1, 0, _put_varint(s)) for s in shape))
    out = _proto_field(1, 0, _put_varint(dtype))
    out += _proto_field(2, 2, _put_varint(len(dims)) + dims)
    if shard_id:
        out += _proto_field(3, 0, _put_varint(shard_id))
    if offset:
        out += _proto_field(4, 0, _put_varint(offset))
    out += _proto_field(5, 0, _put_varint(size))
    out += _proto_field(6, 5, struct.pack("<I", crc))
    return out


class _BlockBuilder(object):
    def __init__(self, restart_interval):
        self.buf = bytearray()
        self.restarts = [0]
        self.counter = 0
        self.last_key = b""
        self.interval = restart_interval

    def add(self, key, value):
        shared = 0
        if self.counter < self.interval:
            m = min(len(self.last_key), len(key))
            while shared < m and self.last_key[shared] == key[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.counter = 0
        self.buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value))
        self.buf += key[shared:] + value
        self.last_key = key
        self.counter += 1

    def finish(self):
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))


def _emit_block(out, contents):
    handle = (len(out), len(contents))
    out += contents + b"\x00"
    out += struct.pack("<I", mask_crc(crc32c(contents + b"\x00")))
    return handle


def write_bundle(prefix, tensors, block_size=4096, restart_interval=16):
    """Write {name: ndarray} as a single-shard V2 bundle (index table + data file)."""
    prefix = str(prefix)
    names = sorted(tensors, key=lambda s: s.encode("utf-8"))
    data = bytearray()
    items = [(b"", _proto_field(1, 0, _put_varint(1)) + _proto_field(3, 2, _put_varint(2) + _proto_field(1, 0, _put_varint(1))))]
    rev = {v: k for k, v in _NP_DTYPES.items()}
    for name in names:
        a = np.ascontiguousarray(tensors[name])
        raw = a.astype(a.dtype.newbyteorder("<")).tobytes()
        items.append((name.encode("utf-8"), _entry_proto(rev[a.dtype.type], a.shape, 0, len(data), len(raw), mask_crc(crc32c(raw)))))
        data += raw
    with open(_shard_path(prefix, 0, 1), "wb") as f:
        f.write(bytes(data))
    out = bytearray()
    index = _BlockBuilder(1)
    block = _BlockBuilder(restart_interval)
    pending = None
    for key, value in items:
        if pending is not None:
            index.add(pending[0], pending[1])
            pending = None
        block.add(key, value)
        if len(block.buf) >= block_size:
            h = _emit_block(out, block.finish())
            pending = (key, _put_varint(h[0]) + _put_varint(h[1]))
            block = _BlockBuilder(restart_interval)
    if block.buf:
        h = _emit_block(out, block.finish())
        pending = (block.last_key, _put_varint(h[0]) + _put_varint(h[1]))
    if pending is not None:
        index.add(pending[0], pending[1])
    meta = _emit_block(out, _BlockBuilder(1).finish())
    idx = _emit_block(out, index.finish())
    footer = _put_varint(meta[0]) + _put_varint(meta[1]) + _put_varint(idx[0]) + _put_varint(idx[1])
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    out += footer
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))
