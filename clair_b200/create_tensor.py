"""CreateTensor with the pileup counting on the GPU.

Mirrors the reference's dataPrepScripts/CreateTensor.py: `OutputAlnTensor(args)` (:179-394) keeps its arguments, child
processes (`samtools faidx`, `samtools view`, `gzip -fdc candidates`) and output rows (:57-62).  What the reference does
per SAM row before the CIGAR walk stays on the host - header rows (:253), the mapping-quality filter (:264) and the
per-POS depth cap (:274-281) - then the surviving reads are encoded into flat arrays (`encode_alignments`) and the walk +
`generate_tensor` (:283-366, :29-65) run as one CUDA kernel over all candidate sites (`clairb_create_tensors`,
csrc/create_tensor_kernels.cuh).  `created_tensor_generator_from` / `TensorBlock.predict` keep the tensors on the device
and hand them straight to the forward pass, which removes the text hop between CreateTensor.py and call_var.py
(clair/callVarBam.py:191-200) and the host->device copy.

Not modelled: the 5,000,000-record memory guard (`available_slots`, :180,285-286,306-308), which only drops records when
more than five million are outstanding - which ones then depends on the iteration order of a Python set (:305).  Reads must be coordinate-sorted (as `samtools view` of a sorted BAM yields them; the
reference's flush at :368-381 assumes the same) and a read whose CIGAR consumes more bases than SEQ holds is an error
(the reference raises IndexError on it once a window is open).
"""
import ctypes
import shlex
import sys

import numpy as np

from . import _lib, param
from .utils import IUPAC_BASES

FLANK = param.flankingBaseNum                    # shared/param.py:9
N_POS = 2 * FLANK + 1
OP_M, OP_I, OP_D = 0, 1, 2
CT_LEFT_EDGE, CT_SUBTRACT = 1, 2

_CODE = np.full(256, -1, np.int8)                # CIGAR char -> kept op code
for _ch in "M=X":
    _CODE[ord(_ch)] = OP_M
_CODE[ord("I")] = OP_I
_CODE[ord("D")] = OP_D
_IS_IUPAC = np.zeros(256, bool)
for _ch in IUPAC_BASES:
    _IS_IUPAC[ord(_ch)] = True
_REF_ADV = np.zeros(256, bool)
_QRY_ADV = np.zeros(256, bool)
for _ch in "M=XD":                               # CreateTensor.py:294,337: ops that advance the reference position
    _REF_ADV[ord(_ch)] = True
for _ch in "SM=XI":                              # :290,294,323: ops that advance the query position
    _QRY_ADV[ord(_ch)] = True


class Alignments(object):
    """Reads of one region, encoded for clairb_create_tensors (field meanings: include/clair_b200.h)."""
    __slots__ = ("read_pos", "read_end", "read_op0", "read_strand", "op_ref", "op_qry", "op_len", "seq")

    @property
    def n_reads(self):
        return int(self.read_pos.shape[0])

    @property
    def n_ops(self):
        return int(self.op_ref.shape[0])


def _pinned_empty(n, dtype):
    """numpy array over page-locked host memory (clairb_host_alloc): uploads from it run at full PCIe rate and
    asynchronously; freed when the array is collected."""
    import weakref
    lib = _lib.load()
    nbytes = max(1, int(n)) * np.dtype(dtype).itemsize
    ptr = ctypes.c_void_p()
    _lib.check(lib.clairb_host_alloc(ctypes.byref(ptr), nbytes), None, "clairb_host_alloc")
    buf = (ctypes.c_char * nbytes).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(n))
    weakref.finalize(buf, lib.clairb_host_free, ptr)
    return arr


def _as_text_block(sam):
    if isinstance(sam, (bytes, bytearray, memoryview)):
        return bytes(sam)
    if isinstance(sam, str):
        return sam.encode("ascii", "replace")
    rows = [r if isinstance(r, bytes) else r.encode("ascii", "replace") for r in sam]
    return b"".join(r if r.endswith(b"\n") else r + b"\n" for r in rows)


def encode_alignments(sam, min_mq=0, dcov=250, encoder="native", pinned=False):
    """`samtools view` rows (an iterable of rows, or one text block) -> Alignments.

    encoder="native": clairb_encode_sam (csrc/encode_host.cuh), multi-threaded host code; with pinned=True the arrays live
    in page-locked memory.  encoder="python": the numpy restatement below, kept as the cross-check."""
    if encoder == "python":
        rows = sam.splitlines() if isinstance(sam, (str, bytes, bytearray)) else sam
        return _encode_alignments_python(rows, min_mq, dcov)
    if encoder != "native":
        raise ValueError("encoder must be 'native' or 'python'")
    lib = _lib.load()
    text = _as_text_block(sam)
    state = (ctypes.c_int32 * 3)(0, 0, -2 ** 31)
    n = [ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()]
    refs = [ctypes.byref(v) for v in n]
    rc = lib.clairb_encode_sam(text, len(text), int(min_mq), int(dcov), state, 0, 0, 0, None, None, None, None, None, None, None,
                               None, *refs)
    _lib.check(rc, None, "clairb_encode_sam")
    R, O, B = (int(v.value) for v in n)
    empty = _pinned_empty if pinned else (lambda k, dt: np.empty(int(k), dt))
    a = Alignments()
    a.read_pos, a.read_end, a.read_op0 = empty(R, np.int32), empty(R, np.int32), empty(R + 1, np.int32)
    a.read_strand = empty(R, np.uint8)
    a.op_ref, a.op_qry, a.op_len = empty(O, np.int32), empty(O, np.int32), empty(O, np.int32)
    a.seq = empty(B, np.uint8)
    ptr = lambda arr: arr.ctypes.data_as(ctypes.c_void_p)
    rc = lib.clairb_encode_sam(text, len(text), int(min_mq), int(dcov), state, R, O, B, ptr(a.read_pos), ptr(a.read_end),
                               ptr(a.read_op0), ptr(a.read_strand), ptr(a.op_ref), ptr(a.op_qry), ptr(a.op_len), ptr(a.seq), *refs)
    _lib.check(rc, None, "clairb_encode_sam")
    return a


def _encode_alignments_python(sam_lines, min_mq=0, dcov=250):
    """The row filters are the reference's (CreateTensor.py:253-281); the CIGAR strings of all kept reads are parsed in one
    vectorised numpy pass."""
    pos_l, strand_l, cigars, seqs = [], [], [], []
    previous_position, depth_cap = 0, 0
    for line in sam_lines:
        if isinstance(line, bytes):
            line = line.decode("ascii", "replace")
        col = line.split()
        if col[0][0] == "@":
            continue
        if int(col[4]) < min_mq:
            continue
        pos = int(col[3]) - 1
        if previous_position != pos:
            previous_position, depth_cap = pos, 0
        else:
            depth_cap += 1
            if depth_cap >= dcov:
                continue
        pos_l.append(pos)
        strand_l.append((int(col[1]) & 16) == 16)
        cigars.append(col[5])
        seqs.append(col[9])

    a = Alignments()
    R = len(pos_l)
    pos = np.asarray(pos_l, np.int64)
    if R and (np.diff(pos) < 0).any():
        raise ValueError("alignments are not coordinate-sorted")
    seq_len = np.fromiter((len(s) for s in seqs), np.int64, R)
    seq_off = np.concatenate([[0], np.cumsum(seq_len)])
    a.seq = np.frombuffer("".join(seqs).encode("ascii", "replace"), np.uint8) if R else np.zeros(0, np.uint8)
    a.read_strand = np.asarray(strand_l, np.uint8)

    # every non-digit character is an op (the digits before it are its length, :287-289); '\n' closes a read
    b = np.frombuffer(("\n".join(cigars) + "\n").encode("ascii", "replace"), np.uint8) if R else np.zeros(0, np.uint8)
    is_digit = (b >= 48) & (b <= 57)
    op_at = np.flatnonzero(~is_digit)
    digit_at = np.flatnonzero(is_digit)
    op_of_digit = np.searchsorted(op_at, digit_at)
    place = op_at[op_of_digit] - digit_at - 1
    if place.size and place.max() > 9:
        raise ValueError("CIGAR op length with more than 10 digits")
    length = np.bincount(op_of_digit, weights=(b[digit_at] - 48) * 10.0 ** place, minlength=op_at.size).astype(np.int64)
    ch = b[op_at]
    newline = ch == 10
    read_of_op = np.cumsum(newline) - newline
    ref_adv = np.where(_REF_ADV[ch], length, 0)
    qry_adv = np.where(_QRY_ADV[ch], length, 0)
    ref_before = np.cumsum(ref_adv) - ref_adv
    qry_before = np.cumsum(qry_adv) - qry_adv
    first_op = np.concatenate([[0], np.flatnonzero(newline)[:-1] + 1]) if R else np.zeros(0, np.int64)
    ref_rel = ref_before - ref_before[first_op][read_of_op]
    qry_rel = qry_before - qry_before[first_op][read_of_op]
    keep = (_CODE[ch] >= 0) & (length > 0)
    reads_base = (keep & (_CODE[ch] != OP_D))
    if (qry_rel[reads_base] + length[reads_base] > seq_len[read_of_op[reads_base]]).any():
        bad = read_of_op[reads_base][np.argmax(qry_rel[reads_base] + length[reads_base] > seq_len[read_of_op[reads_base]])]
        raise ValueError("read %d: CIGAR consumes more bases than SEQ holds" % bad)
    ref_total = np.bincount(read_of_op, weights=ref_adv, minlength=R).astype(np.int64)[:R] if R else np.zeros(0, np.int64)
    end = pos + ref_total
    op_ref = pos[read_of_op[keep]] + ref_rel[keep]
    op_qry = seq_off[read_of_op[keep]] + qry_rel[keep]
    if (R and end.max() >= 2 ** 31 - 64) or a.seq.size >= 2 ** 31 - 64 or (length[keep] >= 2 ** 29).any():
        raise ValueError("block too large for int32 offsets: split the region")
    a.read_pos = pos.astype(np.int32)
    a.read_end = end.astype(np.int32)
    a.read_op0 = np.concatenate([[0], np.cumsum(np.bincount(read_of_op[keep], minlength=R)[:R])]).astype(np.int32)
    a.op_ref = op_ref.astype(np.int32)
    a.op_qry = op_qry.astype(np.int32)
    a.op_len = ((length[keep] << 2) | _CODE[ch[keep]].astype(np.int64)).astype(np.int32)
    return a


class _CAlignments(ctypes.Structure):            # struct clairb_alignments, include/clair_b200.h
    _fields_ = [("read_pos", ctypes.c_void_p), ("read_end", ctypes.c_void_p), ("read_op0", ctypes.c_void_p),
                ("read_strand", ctypes.c_void_p), ("n_reads", ctypes.c_int64),
                ("op_ref", ctypes.c_void_p), ("op_qry", ctypes.c_void_p), ("op_len", ctypes.c_void_p),
                ("n_ops", ctypes.c_int64), ("seq", ctypes.c_void_p), ("seq_len", ctypes.c_int64),
                ("ref", ctypes.c_void_p), ("ref_len", ctypes.c_int64), ("ref_start0", ctypes.c_int32)]


def _ptr(arr):
    return arr.ctypes.data if arr.size else None


class TensorBlock(object):
    """Result of one `create_tensors` call: the sites the reference would have printed, in its output order.

    positions [n] (1-based centres), sequences [n] (the 33-base reference windows of the rows, :59; built on first use),
    depth [n] (aligned bases at the centre), rows [n] (index into the device-resident block), and - when fetched -
    x [n,33,8,4] int16."""

    def __init__(self, model, ctg_name, positions, reference_sequence, window_start, depth, rows, x, subtracted):
        self.model, self.ctg_name = model, ctg_name
        self.positions, self.depth, self.rows, self.x = positions, depth, rows, x
        self.subtracted = subtracted
        self._reference, self._start, self._sequences = reference_sequence, window_start, None
        self._generation = getattr(model, "_ct_generation", 0)      # the device holds one block per handle: the latest

    @property
    def sequences(self):
        if self._sequences is None:
            ref = self._reference
            self._sequences = [ref[s:s + N_POS] for s in self._start.tolist()]                # CreateTensor.py:59
        return self._sequences

    def __len__(self):
        return int(self.positions.shape[0])

    def text(self):
        """The rows CreateTensor.py prints (:57-62) as one bytes block, formatted by the native library
        (clairb_format_tensor_rows); needs raw counts fetched to the host."""
        if self.x is None or self.subtracted:
            raise ValueError("text rows need create_tensors(..., fetch=True, subtract=False)")
        n = len(self)
        if n == 0:
            return b""
        lib = _lib.load()
        ref = self._reference.encode("ascii", "replace")
        positions = np.ascontiguousarray(self.positions, np.int64)
        start = np.ascontiguousarray(self._start, np.int64)
        x = np.ascontiguousarray(self.x, np.int16)
        need = ctypes.c_int64()
        args = (self.ctg_name.encode(), positions.ctypes.data_as(ctypes.c_void_p), ref, len(ref), start.ctypes.data_as(ctypes.c_void_p),
                x.ctypes.data_as(ctypes.c_void_p), n)
        _lib.check(lib.clairb_format_tensor_rows(*args, None, 0, ctypes.byref(need)), None, "clairb_format_tensor_rows")
        buf = ctypes.create_string_buffer(need.value)
        _lib.check(lib.clairb_format_tensor_rows(*args, buf, need.value, ctypes.byref(need)), None, "clairb_format_tensor_rows")
        return buf.raw[:need.value]

    def text_rows(self):
        """The same rows as a list of str (without the newlines)."""
        return self.text().decode("ascii").splitlines()

    def callable_sites(self):
        """Indices (into this block) of the sites tensor_generator_from keeps: centre base is an IUPAC code
        (clair/utils.py:90; a window cut short by the end of the contig fails the same test there by IndexError)."""
        ref = np.frombuffer(self._reference.encode("ascii", "replace"), np.uint8)
        at = self._start + FLANK
        inside = at < ref.shape[0]
        ok = np.zeros(len(self), bool)
        ok[inside] = _IS_IUPAC[ref[at[inside]]]
        return np.flatnonzero(ok).astype(np.int64)

    def predict(self, which=None):
        """Forward pass over sites of this block without the tensors leaving the device -> [n,90] float32."""
        if not self.subtracted:
            raise ValueError("the forward pass takes channel-subtracted tensors: create_tensors(..., subtract=True)")
        which = np.arange(len(self), dtype=np.int64) if which is None else np.asarray(which, np.int64)
        rows = np.ascontiguousarray(self.rows[which])
        out = np.empty((rows.shape[0], _lib.N_OUT), np.float32)
        m = self.model
        with m._lock:
            # checked under the lock: a create_tensors call on another thread replaces the resident block under the same lock
            if getattr(m, "_ct_generation", 0) != self._generation:
                raise RuntimeError("this TensorBlock is no longer resident: a later create_tensors call on the same model replaced it")
            for s in range(0, rows.shape[0], m.max_sites):
                k = min(m.max_sites, rows.shape[0] - s)
                rc = m._lib.clairb_predict_created(m._h, rows[s:s + k].ctypes.data_as(ctypes.c_void_p), k,
                                                   out[s:s + k].ctypes.data_as(ctypes.c_void_p))
                _lib.check(rc, m._h, "clairb_predict_created")
        return out


def create_tensors(model, alignments, candidate_positions, reference_sequence, reference_start_0_based=0, ctg_name="chr",
                   min_coverage=0, consider_left_edge=True, ctg_start=None, ctg_end=None, subtract=False, fetch=True):
    """Counts for every candidate site of one region -> TensorBlock.  `model` is a clair_b200.model.Clair (it owns the
    device handle; weights are only needed for TensorBlock.predict)."""
    if not isinstance(candidate_positions, np.ndarray):
        candidate_positions = list(candidate_positions)
    cand = np.unique(np.asarray(candidate_positions, np.int64))                              # ascending, duplicates are no-ops (:304)
    if ctg_start is not None and ctg_end is not None:
        cand = cand[(cand >= ctg_start) & (cand <= ctg_end)]                                # :83
    cand = cand[cand - reference_start_0_based - (FLANK + 1) >= 0]                          # :55
    ref = np.frombuffer(reference_sequence.encode("ascii", "replace"), np.uint8)
    n = int(cand.shape[0])
    empty = TensorBlock(model, ctg_name, np.zeros(0, np.int64), reference_sequence, np.zeros(0, np.int64), np.zeros(0, np.int32),
                        np.zeros(0, np.int64), np.zeros((0, N_POS, 8, 4), np.int16) if fetch else None, subtract)
    if n == 0 or alignments.n_reads == 0 or ref.size == 0:
        return empty
    if cand.max() >= 2 ** 31 - 64:
        raise ValueError("candidate position beyond int32")
    a = alignments
    ca = _CAlignments(_ptr(a.read_pos), _ptr(a.read_end), _ptr(a.read_op0), _ptr(a.read_strand), a.n_reads,
                      _ptr(a.op_ref), _ptr(a.op_qry), _ptr(a.op_len), a.n_ops, _ptr(a.seq), int(a.seq.size),
                      _ptr(ref), int(ref.size), int(reference_start_0_based))
    centers = cand.astype(np.int32)
    meta = np.empty((n, 2), np.int32)
    x = np.empty((n, N_POS, 8, 4), np.int16) if fetch else None
    flags = (CT_LEFT_EDGE if consider_left_edge else 0) | (CT_SUBTRACT if subtract else 0)
    with model._lock:
        model._ct_generation = getattr(model, "_ct_generation", 0) + 1
        rc = model._lib.clairb_create_tensors(model._h, ctypes.byref(ca), centers.ctypes.data_as(ctypes.c_void_p), n, flags,
                                              x.ctypes.data_as(ctypes.c_void_p) if fetch else None,
                                              meta.ctypes.data_as(ctypes.c_void_p))
        _lib.check(rc, model._h, "clairb_create_tensors")
    rows = np.flatnonzero((meta[:, 0] > 0) & (meta[:, 1] >= min_coverage))                  # a row exists (:303) and :55
    positions = cand[rows]
    start = positions - reference_start_0_based - (FLANK + 1)
    return TensorBlock(model, ctg_name, positions, reference_sequence, start, meta[rows, 1].copy(), rows.astype(np.int64),
                       (x if rows.shape[0] == n else x[rows]) if fetch else None, subtract)


class DeviceTensors(object):
    """Stand-in for the X array of a predict-batch whose tensors live on the device: `Clair.predict` accepts it where
    tensor_generator_from would have handed a [n,33,8,4] array (clair/call_var.py:1343)."""
    __slots__ = ("block", "index")

    def __init__(self, block, index):
        self.block, self.index = block, index

    @property
    def shape(self):
        return (int(self.index.shape[0]), N_POS, param.matrixRow, param.matrixNum)

    def __len__(self):
        return int(self.index.shape[0])

    def _host_rows(self):
        if self.block.x is None:
            raise ValueError("this batch has no host copy of its tensors: an output stage that reads them (batch_output reads "
                             "x[16] and x[17], clair/call_var.py:1021-1151) needs create_tensors(..., subtract=True, fetch=True)")
        return self.block.x

    def __getitem__(self, i):
        """x of site i as the output stage indexes it ([33,8,4], channel-subtracted counts)."""
        return self._host_rows()[self.index[i]]

    def __iter__(self):
        x = self._host_rows()
        for i in self.index.tolist():
            yield x[i]


def device_tensor_generator_from(block, batch_size):
    """Drop-in for utils.tensor_generator_from (clair/utils.py:72-109) in the batch loop: yields (X, [[ctg, pos, seq]] * n)
    per predict-batch for the sites of `block` the reference generator keeps (:90), X being a DeviceTensors."""
    keep = block.callable_sites()
    sequences, positions = block.sequences, block.positions.tolist()
    for s in range(0, keep.shape[0], batch_size):
        idx = keep[s:s + batch_size]
        yield DeviceTensors(block, idx), [[block.ctg_name, str(positions[i]), sequences[i]] for i in idx.tolist()]


def created_tensor_generator_from(block, batch_size):
    """The same batches with the forward already applied: (probabilities [n,90], [[ctg, pos, seq]] * n)."""
    for X, infos in device_tensor_generator_from(block, batch_size):
        yield block.predict(X.index), infos


# ---- the reference's command (children as in CreateTensor.py:115-176) ---------------------------------------------
def _popen(args, binary=False, **kw):
    from subprocess import PIPE, Popen
    return Popen(args, stdout=PIPE, stderr=sys.stderr, bufsize=8388608, universal_newlines=not binary, **kw)


def reference_sequence_from(samtools, reference_file_path, ctg_name, ctg_start, ctg_end, popen=_popen):
    """-> (upper-cased sequence, reference_start_0_based) of `samtools faidx` (CreateTensor.py:115-159)."""
    reference_start = None
    if ctg_start is not None and ctg_end is not None:
        reference_start = max(1, ctg_start - param.expandReferenceRegion)
        region = "%s:%d-%d" % (ctg_name, reference_start, ctg_end + param.expandReferenceRegion)
    else:
        region = ctg_name
    proc = popen(shlex.split("%s faidx %s %s" % (samtools, reference_file_path, region)))
    rows = [row.rstrip() for row in proc.stdout]
    proc.stdout.close()
    proc.wait()
    if proc.returncode != 0 or len(rows) < 2:
        return "", 0
    return "".join(rows[1:]).upper(), 0 if reference_start is None else reference_start - 1


def OutputAlnTensor(args, model=None, popen=_popen, out=None):
    """The reference's entry point (CreateTensor.py:179-394): same arguments, same rows on `--tensor_fn` (PIPE = stdout).

    Two differences from the reference, both by construction of the one-kernel design: (1) the alignments of the requested
    region are encoded as ONE block (the kernel finds a site's reads by binary search over all of them), so host memory
    grows with the region - about 2.4 bytes per aligned base - where the reference streams rows and flushes finished
    windows; a whole contig is processed the way the reference's own drivers cut it, in --ctgStart/--ctgEnd chunks
    (clair/callVarBamParallel.py:90-119), and a block beyond 2^31 ops or bases is refused with "split the region".
    (2) rows come out in ascending position (candidates are sorted and de-duplicated), which is the order of the
    reference for the sorted candidate files its own pipeline produces (ExtractVariantCandidates.py walks a sorted BAM)."""
    reference_sequence, reference_start_0_based = reference_sequence_from(
        args.samtools, args.ref_fn, args.ctgName, args.ctgStart, args.ctgEnd, popen)
    if not reference_sequence:
        print("Failed to load reference seqeunce. Please check if the provided reference fasta %s and the ctgName %s are correct." % (
            args.ref_fn, args.ctgName), file=sys.stderr)
        sys.exit(1)
    cand_proc = None
    if args.can_fn == "PIPE":
        candidate_rows = sys.stdin
    else:
        cand_proc = popen(shlex.split("gzip -fdc %s" % args.can_fn))
        candidate_rows = cand_proc.stdout
    candidates = [int(row.split(maxsplit=2)[1]) for row in candidate_rows]
    if cand_proc is not None:                            # CreateTensor.py:383-385 closes and waits for its children
        cand_proc.stdout.close()
        cand_proc.wait()
    have_region = args.ctgStart is not None and args.ctgEnd is not None
    region = ("%s:%d-%d" % (args.ctgName, args.ctgStart, args.ctgEnd)) if have_region else args.ctgName
    view_args = shlex.split("%s view -F %d %s %s" % (args.samtools, param.SAMTOOLS_VIEW_FILTER_FLAG, args.bam_fn, region))
    own = model is None
    if own:
        from .model import Clair
        model = Clair()
    try:
        # the alignment rows go to the native encoder as one block of bytes (no per-row Python work), and from there to
        # page-locked arrays when a device handle exists
        view = popen(view_args, binary=True) if popen is _popen else popen(view_args)
        rows = view.stdout.read() if hasattr(view.stdout, "read") else view.stdout
        alignments = encode_alignments(rows, args.minMQ, args.dcov, pinned=getattr(model, "_h", None) is not None)
        view.stdout.close()
        view.wait()
        block = create_tensors(model, alignments, candidates, reference_sequence, reference_start_0_based, args.ctgName,
                               args.minCoverage, not args.stop_consider_left_edge, args.ctgStart, args.ctgEnd)
    finally:
        if own:
            model.close()
    text = block.text()
    if out is not None:
        out.extend(text.decode("ascii").splitlines())
    elif args.tensor_fn == "PIPE":
        sys.stdout.write(text.decode("ascii"))
    else:
        import gzip
        with gzip.open(args.tensor_fn, "wb") as f:
            f.write(text)
    return block
