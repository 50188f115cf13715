"""The reference's binary dataset ("bin") without python-blosc (second half).

A bin (dataPrepScripts/Tensor2Bin.py:16-33) is four pickles in a row: the number of sites, then three lists - tensors,
labels, positions - of Blosc frames, one frame per 500 sites (shared/param.py:12), each frame being
`blosc.pack_array(array, cname='lz4hc', clevel=9, shuffle=blosc.NOSHUFFLE)` (clair/utils.py:47-48) = the pickled numpy array,
LZ4-compressed inside a Blosc1 container.  `evaluate.py:64-82` and `train.py` walk the frames with
`utils.decompress_array` (clair/utils.py:223-262) and hand the tensor batches to `Clair.predict`.

Here: `unpack_array` / `decompress_array` / `dataset_info_from` / `no_of_blosc_blocks_from` keep the reference's names,
arguments and return values; the frame itself is decoded by the C-ABI library (`clairb_blosc_decompress`,
csrc/blosc_host.cuh, host code).  `pack_array` writes frames of the same layout so that bins can be produced and round-tripped
without python-blosc (greedy LZ4, not lz4hc's optimal parse: same format, larger output).

PARITY: `decompress_array`'s block walk is pinned on the reference's own function (tests run it from /root/reference with
this module's `unpack_array` behind its `blosc` import; committed golden: tests/golden/bins_walk.json).  The frame decoder is
UNPINNED: python-blosc 1.8.3 is not vendored with the reference and is absent here, so there is no frame written by the
real library to read; it is tested on frames assembled from the published Blosc1 / LZ4 formats.
"""
import ctypes
import pickle
import struct
from collections import namedtuple

import numpy as np

from . import _lib, param

bloscBlockSize = 500                 # shared/param.py:12
trainingDatasetPercentage = 0.9      # shared/param.py:20

DatasetInfo = namedtuple("DatasetInfo", [                                  # clair/utils.py:264 region
    "dataset_size", "x_array_compressed", "y_array_compressed", "position_array_compressed",
    "no_of_training_examples_from_train_binary", "is_separated_train_and_validation_binary"])


# ---- frames ----------------------------------------------------------------------------------------------------------
def blosc_decompress(frame):
    """One Blosc1 frame -> bytes."""
    lib = _lib.load()
    frame = bytes(frame)
    n = ctypes.c_int64()
    rc = lib.clairb_blosc_decompress(frame, len(frame), None, 0, ctypes.byref(n))
    _lib.check(rc, None, "clairb_blosc_decompress")
    out = ctypes.create_string_buffer(max(1, n.value))
    rc = lib.clairb_blosc_decompress(frame, len(frame), out, n.value, ctypes.byref(n))
    _lib.check(rc, None, "clairb_blosc_decompress")
    return out.raw[:n.value]


def unpack_array(packed_array):
    """blosc.unpack_array: the numpy array of one frame (clair/utils.py:241)."""
    return pickle.loads(blosc_decompress(packed_array))


def _lz4_greedy(data):
    """A valid LZ4 block of `data` (hash of 4 bytes -> last position, greedy matches)."""
    n = len(data)
    out = bytearray()
    table = {}
    anchor = i = 0
    end_of_matches = n - 12                   # the format keeps the last 5 bytes literal and starts no match in the last 12
    while i < end_of_matches:
        key = data[i:i + 4]
        cand = table.get(key)
        table[key] = i
        if cand is None or i - cand > 65535:
            i += 1
            continue
        m = 4
        limit = n - 5
        while i + m < limit and data[cand + m] == data[i + m]:
            m += 1
        lit = i - anchor
        token_l, token_m = min(lit, 15), min(m - 4, 15)
        out.append((token_l << 4) | token_m)
        if lit >= 15:
            rest = lit - 15
            out += b"\xff" * (rest // 255) + bytes([rest % 255])
        out += data[anchor:i]
        out += struct.pack("<H", i - cand)
        if m - 4 >= 15:
            rest = m - 4 - 15
            out += b"\xff" * (rest // 255) + bytes([rest % 255])
        i += m
        anchor = i
    lit = n - anchor
    out.append(min(lit, 15) << 4)
    if lit >= 15:
        rest = lit - 15
        out += b"\xff" * (rest // 255) + bytes([rest % 255])
    out += data[anchor:]
    return bytes(out)


def blosc_compress(data, typesize=1, blocksize=None, split=True, store=False):
    """bytes -> one Blosc1 frame with the LZ4 codec and no shuffle (layout: csrc/blosc_host.cuh).  split=True splits every
    full block into `typesize` streams the way c-blosc 1.x does for LZ4 (its forward-compatible split mode)."""
    data = bytes(data)
    n = len(data)
    typesize = typesize if 0 < typesize <= 255 else 1
    if blocksize is None:
        blocksize = max(typesize * 128, min(n, 1 << 18)) if n else typesize
        blocksize -= blocksize % typesize
    flags = (1 << 5) | (0 if split else 0x10)
    if store or n < 128:
        return struct.pack("<BBBBIII", 2, 1, flags | 0x2, typesize, n, blocksize, 16 + n) + data
    nblocks = (n + blocksize - 1) // blocksize
    body = bytearray()
    starts = []
    for b in range(nblocks):
        block = data[b * blocksize:(b + 1) * blocksize]
        starts.append(16 + 4 * nblocks + len(body))
        leftover = len(block) != blocksize
        nsplits = typesize if (split and typesize <= 16 and blocksize // typesize >= 128 and not leftover) else 1
        ne = len(block) // nsplits
        for s in range(nsplits):
            part = block[s * ne:(s + 1) * ne] if nsplits > 1 else block
            comp = _lz4_greedy(part)
            if len(comp) >= len(part):
                comp = part                       # stored: compressed size == stream size
            body += struct.pack("<i", len(comp)) + comp
    head = struct.pack("<BBBBIII", 2, 1, flags, typesize, n, blocksize, 16 + 4 * nblocks + len(body))
    return head + b"".join(struct.pack("<i", s) for s in starts) + bytes(body)


def pack_array(array, split=True):
    """blosc.pack_array(array, cname='lz4hc', clevel=9, shuffle=NOSHUFFLE) (clair/utils.py:47-48): same container, greedy LZ4."""
    array = np.asarray(array)
    return blosc_compress(pickle.dumps(array, pickle.HIGHEST_PROTOCOL), typesize=array.itemsize, split=split)


# ---- the walk over frames (clair/utils.py:223-262) -------------------------------------------------------------------
def decompress_array(array, blosc_start_index, first_blosc_block_data_index, no_of_data_rows_to_retrieve,
                     no_of_blosc_blocks, read_index_list=None):
    """Rows [.. no_of_data_rows_to_retrieve) starting at row `first_blosc_block_data_index` of frame `blosc_start_index`
    -> (rows, next first-row index, next frame index), with the reference's conventions: a walk that starts inside a frame
    returns only the rest of that frame; (-1, -1) after the last frame; (None, -1, -1) when there is nothing left."""
    def frame(i):
        return unpack_array(array[i if read_index_list is None else read_index_list[i]])

    want = no_of_data_rows_to_retrieve
    if blosc_start_index >= no_of_blosc_blocks:
        return None, -1, -1
    if first_blosc_block_data_index > 0:                  # resume inside a frame: its remaining rows, whatever their number
        return frame(blosc_start_index)[first_blosc_block_data_index:], 0, blosc_start_index + 1
    parts, have, nxt = [], 0, blosc_start_index
    while nxt < no_of_blosc_blocks and have < want:
        parts.append(frame(nxt))
        have += len(parts[-1])
        nxt += 1
    if have <= 0:
        return None, -1, -1
    rows = np.concatenate(parts)
    if have < want:                                       # ran out of frames
        return rows, -1, -1
    extra = have % want                                   # rows of the last frame that belong to the next call
    resume_frame, resume_row = (nxt, 0) if extra == 0 else (nxt - 1, len(parts[-1]) - extra)
    if resume_frame >= no_of_blosc_blocks:
        resume_frame = resume_row = -1
    return rows[:want], resume_row, resume_frame


def dataset_info_from(binary_file_path=None, train_binary_file_path=None, validation_binary_file_path=None):
    """clair/utils.py:265-312 for the binary inputs (the text route builds a bin first: Tensor2Bin.py)."""
    def load(path):
        with open(path, "rb") as fh:
            return pickle.load(fh), pickle.load(fh), pickle.load(fh), pickle.load(fh)

    from_train = None
    if train_binary_file_path is not None and validation_binary_file_path is not None:
        size, xs, ys, ps = load(train_binary_file_path)
        from_train = size
        vsize, vxs, vys, vps = load(validation_binary_file_path)
        size, xs, ys, ps = size + vsize, xs + vxs, ys + vys, ps + vps
    elif binary_file_path is not None:
        size, xs, ys, ps = load(binary_file_path)
    else:
        raise ValueError("a binary file path is required")
    return DatasetInfo(size, xs, ys, ps, from_train, from_train is not None)


def no_of_blosc_blocks_from(dataset_info, no_of_training_examples, blosc_block_size=bloscBlockSize):
    """clair/utils.py:365-376"""
    if dataset_info.is_separated_train_and_validation_binary:
        no_of_validation_examples = dataset_info.dataset_size - no_of_training_examples
        return int(np.ceil(float(no_of_training_examples) / blosc_block_size)) + \
            int(np.ceil(float(no_of_validation_examples) / blosc_block_size))
    return int(np.ceil(float(dataset_info.dataset_size) / blosc_block_size))


def write_bin(path, X, Y, positions, block_size=bloscBlockSize):
    """Tensor2Bin.py:26-33 from arrays already in memory: X [n,33,8,4], Y [n,labels], positions [n] strings."""
    n = len(X)
    xs, ys, ps = [], [], []
    for s in range(0, n, block_size):
        xs.append(pack_array(np.asarray(X[s:s + block_size])))
        ys.append(pack_array(np.asarray(Y[s:s + block_size])))
        ps.append(pack_array(np.asarray(positions[s:s + block_size])))
    with open(path, "wb") as fh:
        for obj in (n, xs, ys, ps):
            pickle.dump(obj, fh, protocol=pickle.HIGHEST_PROTOCOL)


def prediction_batches_from(dataset_info, batch_size=None):
    """The loop of evaluate.py:64-86: yields (x_batch, y_batch) of `batch_size` sites (param.predictBatchSize) in file order;
    x_batch goes to Clair.predict as is."""
    batch_size = batch_size or param.predictBatchSize
    no_of_training_examples = dataset_info.no_of_training_examples_from_train_binary or \
        int(dataset_info.dataset_size * trainingDatasetPercentage)
    no_of_blocks = no_of_blosc_blocks_from(dataset_info, no_of_training_examples)
    blosc_index, first = 0, 0
    while blosc_index != -1:
        x_batch, next_first, next_index = decompress_array(dataset_info.x_array_compressed, blosc_index, first, batch_size, no_of_blocks)
        y_batch, _, _ = decompress_array(dataset_info.y_array_compressed, blosc_index, first, batch_size, no_of_blocks)
        if x_batch is None:
            return
        yield x_batch, y_batch
        blosc_index, first = next_index, next_first
