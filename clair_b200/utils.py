"""Tensor text decode feeding the forward path: mirror of the reference generator
``clair.utils.tensor_generator_from`` (reference clair/utils.py:72-109, batches_from :55-65).

Wire format (CreateTensor.py:60-65): one site per line, whitespace separated:
``ctg  pos  seq33  v_0 ... v_1055`` with the 1056 integers in [33][8][4] order.
Behaviour kept from the reference: rows whose centre base seq[16] is not an IUPAC code are
dropped (utils.py:90); channels 1..3 have channel 0 subtracted (utils.py:96-98); a progress line
goes to stderr per batch (utils.py:101); empty batches are skipped (utils.py:103-104); the
yielded X is a float32 view of the first n rows (utils.py:105).
"""
import shlex
import sys
from subprocess import PIPE, Popen

import numpy as np

from . import param

IUPAC_BASES = frozenset("ACGTURYSWKMBDHVN")            # keys of shared/utils.py:19-29
no_of_positions, matrix_row, matrix_num = param.no_of_positions, param.matrixRow, param.matrixNum
input_tensor_size = param.input_tensor_size


def subtract_reference_channel(X):
    """In place: X[..., i] -= X[..., 0] for i = 1..3 (clair/utils.py:97-98)."""
    X[..., 1:] -= X[..., 0:1]
    return X


def rows_to_batch(rows, batch_size, out=None):
    """Decode up to batch_size text rows -> (X[:n] float32 [n,33,8,4], non_tensor_infos[:n])."""
    if out is None:
        out = np.empty((batch_size, input_tensor_size), dtype=np.float32)
    infos = []
    for row in rows:
        columns = row.split()
        info = columns[:-input_tensor_size]
        _, _, sequence = info
        if sequence[param.flankingBaseNum] not in IUPAC_BASES:
            continue
        out[len(infos)] = np.array(columns[-input_tensor_size:], dtype=np.float32)
        infos.append(info)
    n = len(infos)
    X = out.reshape((batch_size, no_of_positions, matrix_row, matrix_num))
    subtract_reference_channel(X[:n])
    return X[:n], infos


def native_rows_to_batch(lines, batch_size, out=None, dtype=np.float32):
    """Same contract as rows_to_batch for one batch of text lines (bytes or str), decoded by the C-ABI library's
    clairb_decode_rows (one call per batch instead of a Python split + float parse per row)."""
    import ctypes
    from . import _lib
    lib = _lib.load()
    dtype = np.dtype(dtype)
    code = _lib.DTYPE_I16 if dtype == np.int16 else _lib.DTYPE_F32
    if len(lines) > batch_size:
        raise ValueError("%d rows do not fit a batch of %d" % (len(lines), batch_size))
    if out is None:
        out = np.empty((batch_size, input_tensor_size), dtype=dtype)
    elif out.dtype != dtype or out.size < batch_size * input_tensor_size or not out.flags["C_CONTIGUOUS"]:
        raise ValueError("`out` must be a C-contiguous %s array of at least %d x %d values" % (dtype, batch_size, input_tensor_size))
    text = b"".join((ln if isinstance(ln, bytes) else ln.encode()) if ln[-1:] in (b"\n", "\n")
                    else (ln if isinstance(ln, bytes) else ln.encode()) + b"\n" for ln in lines)
    infos = []
    n = 0
    if lines:
        off = np.empty((len(lines), 6), dtype=np.int32)
        rows_read, rows_kept, consumed = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        rc = lib.clairb_decode_rows(text, len(text), len(lines), code, out.ctypes.data_as(ctypes.c_void_p),
                                    off.ctypes.data_as(ctypes.c_void_p), ctypes.byref(rows_read), ctypes.byref(rows_kept),
                                    ctypes.byref(consumed))
        _lib.check(rc, None, "clairb_decode_rows")
        if rows_read.value != len(lines):
            raise ValueError("clairb_decode_rows read %d of %d rows" % (rows_read.value, len(lines)))
        n = rows_kept.value
        assert n <= len(lines) <= batch_size
        o = off[:n].tolist()
        infos = [[text[a:b].decode(), text[c:d].decode(), text[e:f].decode()] for a, b, c, d, e, f in o]
    X = out.reshape(-1)[:batch_size * input_tensor_size].reshape((batch_size, no_of_positions, matrix_row, matrix_num))
    return X[:n], infos


def tensor_generator_from(tensor_file_path, batch_size, alloc=None, decoder="native", dtype=np.float32):
    """Yield (X, non_tensor_infos) per batch.  tensor_file_path "PIPE" reads stdin, anything else
    goes through ``gzip -fdc`` like the reference.  `alloc(shape, dtype)` lets the caller hand out
    pinned buffers (clair_b200.model.pinned_empty) so predict()'s H2D copy is asynchronous.
    decoder: "native" = clairb_decode_rows of the C-ABI library (default), "python" = the row-by-row mirror of the
    reference code (kept as the cross-check).  dtype: float32 like the reference, or int16 (native decoder only:
    the same counts as the compact CLAIRB_DTYPE_I16 transport)."""
    if decoder not in ("native", "python"):
        raise ValueError("decoder must be 'native' or 'python'")
    if decoder == "python" and np.dtype(dtype) != np.float32:
        raise ValueError("the python decoder yields float32 only")
    proc = None
    if tensor_file_path != "PIPE":
        proc = Popen(shlex.split("gzip -fdc %s" % (tensor_file_path)), stdout=PIPE, bufsize=8388608,
                     universal_newlines=(decoder == "python"))
        fo = proc.stdout
    else:
        fo = sys.stdin if decoder == "python" else getattr(sys.stdin, "buffer", sys.stdin)

    processed_tensors = 0
    it = iter(fo)
    exhausted = False
    while not exhausted:
        rows = []
        for _ in range(batch_size):
            try:
                rows.append(next(it))
            except StopIteration:
                exhausted = True
                break
        buf = alloc((batch_size, input_tensor_size), np.dtype(dtype)) if alloc is not None else None
        if decoder == "native":
            X, infos = native_rows_to_batch(rows, batch_size, out=buf, dtype=dtype)
        else:
            X, infos = rows_to_batch(rows, batch_size, out=buf)
        processed_tensors += len(infos)
        print("Processed %d tensors" % processed_tensors, file=sys.stderr)
        if len(infos) <= 0:
            continue
        yield X, infos

    if proc is not None:
        fo.close()
        proc.wait()


def format_tensor_row(ctg, pos, seq, counts):
    """Inverse of the decode: the text row CreateTensor prints (CreateTensor.py:60-65)."""
    flat = np.asarray(counts).reshape(-1)
    return "%s %d %s %s" % (ctg, pos, seq, " ".join("%d" % v for v in flat))


# ---- binary tensor transport --------------------------------------------------------------------
# The text wire format costs ~2.5 KB and a str.split + float parse per site (7 k rows/s/thread in the reference).  The
# binary framing carries exactly the same fields - contig, position, the 33-base sequence and the 1056 raw integer
# counts of CreateTensor.py:60-65 - as fixed-size little-endian records, so a batch is one read() and one numpy view:
#     file  = MAGIC (8 bytes) | record*
#     record = ctg[32] (NUL padded) | pos int64 | seq[33] | 7 pad bytes | counts int16[33][8][4]      = 2192 bytes
# Decoding applies what tensor_generator_from applies to text rows: the IUPAC filter on seq[16] (utils.py:90) and
# X[..., 1:] -= X[..., 0:1] (utils.py:96-98), here in int16 (exact: the counts are small non-negative integers), and
# yields X as int16 - the CLAIRB_DTYPE_I16 transport predict() accepts, half the host->device bytes of float32.
BINARY_MAGIC = b"CLRBT\x01\x00\n"
BINARY_RECORD = np.dtype([("ctg", "S32"), ("pos", "<i8"), ("seq", "S33"), ("pad", "V7"),
                          ("counts", "<i2", (no_of_positions, matrix_row, matrix_num))])
assert BINARY_RECORD.itemsize == 2192


def write_binary_tensors(fo, sites):
    """Write (ctg, pos, seq, counts[33,8,4]) tuples to the binary file object `fo` (MAGIC first)."""
    fo.write(BINARY_MAGIC)
    rec = np.zeros(1, dtype=BINARY_RECORD)
    for ctg, pos, seq, counts in sites:
        ctg_b, seq_b = str(ctg).encode(), str(seq).encode()
        if len(ctg_b) > 32 or len(seq_b) != 2 * param.flankingBaseNum + 1:
            raise ValueError("contig name longer than 32 bytes or sequence not %d bases" % (2 * param.flankingBaseNum + 1))
        c = np.asarray(counts).reshape(no_of_positions, matrix_row, matrix_num)
        if c.min() < -32768 or c.max() > 32767:
            raise ValueError("count outside the int16 range")
        rec["ctg"], rec["pos"], rec["seq"], rec["counts"] = ctg_b, int(pos), seq_b, c
        fo.write(rec.tobytes())


def text_to_binary(text_rows, fo):
    """Convert CreateTensor text rows (an iterable of str) to the binary framing."""
    def sites():
        for row in text_rows:
            columns = row.split()
            if not columns:
                continue
            ctg, pos, seq = columns[:-input_tensor_size]
            yield ctg, int(pos), seq, np.array(columns[-input_tensor_size:], dtype=np.int64)
    write_binary_tensors(fo, sites())


def binary_tensor_generator_from(tensor_file_path, batch_size, alloc=None, as_float32=False):
    """Binary counterpart of tensor_generator_from: yields (X, non_tensor_infos) per batch with the same filtering,
    transform, progress lines and info triples ([ctg, str(pos), seq]).  X is int16 (or float32 when asked);
    `alloc(shape, dtype)` lets the caller hand out pinned buffers."""
    fo = sys.stdin.buffer if tensor_file_path == "PIPE" else open(tensor_file_path, "rb")
    try:
        if fo.read(len(BINARY_MAGIC)) != BINARY_MAGIC:
            raise ValueError("%s: not a clair_b200 binary tensor stream" % tensor_file_path)
        processed_tensors = 0
        out_dtype = np.float32 if as_float32 else np.int16
        while True:
            raw = fo.read(batch_size * BINARY_RECORD.itemsize)
            if len(raw) % BINARY_RECORD.itemsize:
                raise ValueError("truncated binary tensor record")
            if not raw:
                break
            recs = np.frombuffer(raw, dtype=BINARY_RECORD)
            centre = np.frombuffer(raw, dtype=np.uint8).reshape(len(recs), BINARY_RECORD.itemsize)[
                :, BINARY_RECORD.fields["seq"][1] + param.flankingBaseNum]
            keep = np.isin(centre, np.frombuffer("".join(sorted(IUPAC_BASES)).encode(), dtype=np.uint8))
            recs = recs[keep]
            n = len(recs)
            processed_tensors += n
            print("Processed %d tensors" % processed_tensors, file=sys.stderr)
            if n == 0:
                continue
            shape = (batch_size, no_of_positions, matrix_row, matrix_num)
            X = alloc(shape, out_dtype) if alloc is not None else np.empty(shape, dtype=out_dtype)
            X[:n] = recs["counts"]
            subtract_reference_channel(X[:n])
            infos = [[r["ctg"].decode(), str(int(r["pos"])), r["seq"].decode()] for r in recs]
            yield X[:n], infos
            if len(raw) < batch_size * BINARY_RECORD.itemsize:
                break
    finally:
        if fo is not sys.stdin.buffer:
            fo.close()
