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


def tensor_generator_from(tensor_file_path, batch_size, alloc=None):
    """Yield (X, non_tensor_infos) per batch.  tensor_file_path "PIPE" reads stdin, anything else
    goes through ``gzip -fdc`` like the reference.  `alloc(shape, dtype)` lets the caller hand out
    pinned buffers (clair_b200.model.pinned_empty) so predict()'s H2D copy is asynchronous."""
    proc = None
    if tensor_file_path != "PIPE":
        proc = Popen(shlex.split("gzip -fdc %s" % (tensor_file_path)), stdout=PIPE, bufsize=8388608,
                     universal_newlines=True)
        fo = proc.stdout
    else:
        fo = sys.stdin

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
        buf = alloc((batch_size, input_tensor_size), np.float32) if alloc is not None else None
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
