"""CreateTensor on the device.  (The file name sorts after the forward-path tests on purpose: the
driver runs `pytest -x`, and a widening row must not be able to hide the main path.)

not gpu: the oracle restatement and the kernel's per-site rule (its __host__ __device__ half compiled for the CPU by
         tests/harness) against rows printed by the reference's own OutputAlnTensor (tests/golden/create_tensor_cases.json.gz,
         made by oracle/gen_golden_create_tensor.py); the host encoder.
gpu:     the CUDA kernel through the C-ABI against the same golden rows and, on larger random regions, against the oracle;
         tensors born on the device fed to the forward pass against the host-fed forward (bit-identical).
"""
import ctypes
import gzip
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clair_b200 import create_tensor as CT                      # noqa: E402
from oracle import create_tensor_oracle as O                    # noqa: E402
from oracle import gen_golden_create_tensor as G                # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "create_tensor_cases.json.gz")


def golden_cases():
    with gzip.open(GOLDEN) as f:
        return json.loads(f.read())


def reference_window(case):
    """What `samtools faidx` would have handed the reference for this case (CreateTensor.py:115-159)."""
    a = case["args"]
    if a["ctgStart"] is None:
        return case["contig"].upper(), 0
    lo = max(1, a["ctgStart"] - a["expandReferenceRegion"])
    hi = a["ctgEnd"] + a["expandReferenceRegion"]
    return case["contig"][lo - 1:hi].upper(), lo - 1


def candidates_of(case):
    return [int(l.split()[1]) for l in case["candidates"]]


def oracle_rows(case):
    a = case["args"]
    seq, start0 = reference_window(case)
    return O.create_tensors(case["sam"], candidates_of(case), seq, start0, a["ctgName"], a["minMQ"], a["dcov"],
                            a["minCoverage"], not a["stop_consider_left_edge"], a["ctgStart"], a["ctgEnd"])


CASES = golden_cases()
CASE_IDS = [c["name"] for c in CASES]


@pytest.mark.parametrize("case", CASES, ids=CASE_IDS)
def test_oracle_reproduces_reference_rows(case):
    assert [O.format_row(r) for r in oracle_rows(case)] == case["expected"]


# ---- the kernel's rule, compiled for the host ------------------------------------------------------------------------
WALKS = {"staged": [0, 0]}                      # (site, read) pairs that took the staged / the general walk, over the module


@pytest.fixture(scope="module", params=["staged walk (-DCLAIRB_CT_STAGED)", "op-major walk (the kernel's default)",
                                        "position-major walk (-DCLAIRB_CT_FLAT)"])
def host_rule(request, tmp_path_factory):
    out = str(tmp_path_factory.mktemp("ct_harness") / "ct_host.so")
    src = os.path.join(ROOT, "tests", "harness", "create_tensor_host.cu")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    variant = ["-DCLAIRB_CT_FLAT"] if "FLAT" in request.param else []
    subprocess.run([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC"] + variant +
                   ["-o", out, src], check=True)
    lib = ctypes.CDLL(out)
    lib.ct_host_sites.restype = ctypes.c_int
    lib.ct_host_sites_staged.restype = ctypes.c_int
    lib.staged = request.param.startswith("staged")
    return lib


def host_rule_rows(lib, case):
    """Same host-side steps as clair_b200.create_tensor.create_tensors, the device call replaced by the host-compiled rule."""
    a = case["args"]
    seq, start0 = reference_window(case)
    aln = CT.encode_alignments(case["sam"], a["minMQ"], a["dcov"])
    cand = np.unique(np.asarray(candidates_of(case), np.int64))
    if a["ctgStart"] is not None:
        cand = cand[(cand >= a["ctgStart"]) & (cand <= a["ctgEnd"])]
    cand = cand[cand - start0 - 17 >= 0]
    centers = cand.astype(np.int32)
    n = centers.shape[0]
    counts = np.zeros((n, 1056), np.int32)
    opened = np.zeros(n, np.int32)
    ref = np.frombuffer(seq.encode(), np.uint8)
    maxend = np.maximum.accumulate(aln.read_end) if aln.n_reads else np.zeros(0, np.int32)

    def p(arr):
        return arr.ctypes.data_as(ctypes.c_void_p)

    args = [p(aln.read_pos), p(aln.read_end), p(maxend), p(aln.read_op0), p(aln.read_strand),
            ctypes.c_int32(aln.n_reads), p(aln.op_ref), p(aln.op_qry), p(aln.op_len), p(aln.seq), p(ref),
            ctypes.c_int32(start0), ctypes.c_int32(ref.shape[0]), p(centers), ctypes.c_int32(n),
            ctypes.c_int(0 if a["stop_consider_left_edge"] else 1), p(counts), p(opened)]
    if lib.staged:
        padded = np.zeros(aln.seq.shape[0] + 64, np.uint8)            # the device buffer carries the same slack behind SEQ
        padded[:aln.seq.shape[0]] = aln.seq
        assert padded.ctypes.data % 4 == 0
        args[9] = p(padded)
        took = (ctypes.c_int64 * 2)()
        rc = lib.ct_host_sites_staged(*args, ctypes.byref(took, 0), ctypes.byref(took, 8))
        WALKS["staged"][0] += took[0]
        WALKS["staged"][1] += took[1]
    else:
        rc = lib.ct_host_sites(*args)
    assert rc == 0, "a read outside the searched range opens site %d" % (rc - 1)
    rows = []
    for i in range(n):
        depth = counts[i].reshape(33, 8, 4)[16, :, 0].sum()
        if opened[i] > 0 and depth >= a["minCoverage"]:
            s = int(cand[i]) - start0 - 17
            rows.append((a["ctgName"], int(cand[i]), seq[s:s + 33], counts[i]))
    return rows


@pytest.mark.parametrize("case", CASES, ids=CASE_IDS)
def test_kernel_rule_on_host_reproduces_reference_rows(host_rule, case):
    assert [O.format_row(r) for r in host_rule_rows(host_rule, case)] == case["expected"]


def test_kernel_rule_on_host_random_regions(host_rule):
    rng = np.random.default_rng(77)
    for k in range(6):
        case = G.make_case(rng, "rnd%d" % k, 4000, 900, 400, style=["mixed", "dense", "leading"][k % 3],
                           stop_left=(k == 4), dup_rate=0.2, dcov=5)
        want = oracle_rows(case)
        got = host_rule_rows(host_rule, case)
        assert len(got) == len(want)
        for g, w in zip(got, want):
            assert g[:3] == w[:3] and np.array_equal(g[3], w[3].reshape(-1))


def test_staged_and_general_walks_were_both_taken():
    """Runs after the host-rule tests above (file order): the golden regions and the random ones must have sent reads down both
    walks of the staged variant - windows that fit the stage, and windows with more ops / inserted bases than it holds."""
    staged, general = WALKS["staged"]
    assert staged > 1000 and general > 0, (staged, general)


@pytest.mark.parametrize("shape", [dict(depth=60, read_len=150, ops_per_read=3, site_spacing=40),       # short reads, few ops
                                   dict(depth=25, read_len=6000, ops_per_read=21, site_spacing=60),    # long accurate reads
                                   dict(depth=30, read_len=3000, ops_per_read=601, site_spacing=25)])  # long noisy reads
def test_kernel_rule_on_host_sequencing_shapes(host_rule, shape):
    """The synthetic regions the bench uses (clair_b200.synth.synthetic_alignments) at three read shapes: the kernel's rule
    on the host against the oracle, through SAM text and the native encoder."""
    from clair_b200 import synth
    aln, reference, sites = synth.synthetic_alignments(12000, seed=31, **shape)
    sam = synth.alignments_to_sam(aln, name="syn")
    again = CT.encode_alignments(sam)
    for field in CT.Alignments.__slots__:
        assert np.array_equal(getattr(aln, field), getattr(again, field)), field
    case = {"sam": sam, "contig": reference, "candidates": ["syn\t%d" % p for p in sites],
            "args": {"dcov": 250, "stop_consider_left_edge": False, "minCoverage": 0, "minMQ": 0, "ctgName": "syn",
                     "ctgStart": None, "ctgEnd": None, "expandReferenceRegion": 1000000}}
    want = oracle_rows(case)
    got = host_rule_rows(host_rule, case)
    assert len(want) > 100 and len(got) == len(want)
    for g, w in zip(got, want):
        assert g[:3] == w[:3] and np.array_equal(g[3], w[3].reshape(-1))


# ---- host encoder ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("encoder", ["native", "python"])
def test_encoder_filters_and_offsets(encoder):
    sam = ["@HD\tVN:1.6",
           "r0\t0\tc\t11\t60\t2S3M1I2M2D1M3H\t*\t0\t0\tNNACGTTTA\t*",      # S skips 2 query bases, H skips nothing
           "r1\t16\tc\t11\t5\t4M\t*\t0\t0\tACGT\t*",                         # below minMQ
           "r2\t16\tc\t11\t60\t2M5N2M\t*\t0\t0\tacgt\t*",                    # N advances nothing (CreateTensor.py:283-366)
           "r3\t0\tc\t11\t60\t4M\t*\t0\t0\tACGT\t*",                         # third read at POS 10 -> depth cap 2 drops it
           "r4\t0\tc\t15\t60\t*\t*\t0\t0\tAC\t*"]
    a = CT.encode_alignments(sam, min_mq=10, dcov=2, encoder=encoder)
    assert a.read_pos.tolist() == [10, 10, 14] and a.read_strand.tolist() == [0, 1, 0]
    assert a.read_end.tolist() == [18, 14, 14]
    assert a.read_op0.tolist() == [0, 5, 7, 7]
    assert (a.op_len & 3).tolist() == [0, 1, 0, 2, 0, 0, 0] and (a.op_len >> 2).tolist() == [3, 1, 2, 2, 1, 2, 2]
    assert a.op_ref.tolist() == [10, 13, 13, 15, 17, 10, 12]
    assert a.op_qry.tolist() == [2, 5, 6, 8, 8, 9, 11]
    assert bytes(a.seq) == b"NNACGTTTAacgtAC"
    with pytest.raises(ValueError):                       # CIGAR consumes more bases than SEQ holds
        CT.encode_alignments(["r\t0\tc\t5\t60\t10M\t*\t0\t0\tACGT\t*"], encoder=encoder)
    with pytest.raises(ValueError):                       # not coordinate-sorted
        CT.encode_alignments(["a\t0\tc\t9\t60\t1M\t*\t0\t0\tA\t*", "b\t0\tc\t5\t60\t1M\t*\t0\t0\tA\t*"], encoder=encoder)
    empty = CT.encode_alignments(["@SQ\tSN:c"], encoder=encoder)
    assert empty.n_reads == 0 and empty.n_ops == 0 and empty.read_op0.tolist() == [0]


@pytest.mark.parametrize("case", CASES, ids=CASE_IDS)
def test_native_encoder_equals_python_encoder(case):
    a = case["args"]
    native = CT.encode_alignments(case["sam"], a["minMQ"], a["dcov"])
    python = CT.encode_alignments(case["sam"], a["minMQ"], a["dcov"], encoder="python")
    for field in CT.Alignments.__slots__:
        assert np.array_equal(getattr(native, field), getattr(python, field)), field
    # one text block, CRLF rows, no final newline: same arrays
    block = "\r\n".join(case["sam"])
    again = CT.encode_alignments(block, a["minMQ"], a["dcov"])
    for field in CT.Alignments.__slots__:
        assert np.array_equal(getattr(native, field), getattr(again, field)), field


def test_native_encoder_equals_python_encoder_on_odd_cigars():
    """CIGAR strings the SAM specification does not produce but the reference's character loop defines anyway
    (CreateTensor.py:283-366): unknown op letters, lower case, zero lengths, digits without an op, '*', very long lengths."""
    rng = np.random.default_rng(404)
    alphabet = list("MIDNSHP=XBmid*") + ["", "0"]
    rows, pos = [], 3
    for i in range(300):
        pos += int(rng.integers(0, 4))
        parts = []
        for _ in range(int(rng.integers(0, 9))):
            parts.append("%s%s" % (rng.choice(["", "0", "1", "3", "12", "007"]), rng.choice(alphabet)))
        cigar = "".join(parts) + rng.choice(["", "5", "2M"])
        if not cigar:
            cigar = "*"
        flag = int(rng.choice([0, 16, 2048, 272]))
        sep = rng.choice(["\t", " ", "\t\t"])
        row = sep.join(["q%d" % i, str(flag), "c", str(pos), str(int(rng.integers(0, 61))), cigar, "*", "0", "0",
                        "".join(rng.choice(list("ACGTNacgt"), size=120)), "*", "NM:i:3 extra column"])
        try:                                              # digits run together ("12" + "3M"): both encoders refuse a CIGAR
            CT.encode_alignments([row], encoder="python")  # that consumes more bases than SEQ holds
        except ValueError:
            with pytest.raises(ValueError):
                CT.encode_alignments([row])
            continue
        rows.append(row)
    for min_mq, dcov in ((0, 250), (30, 2)):
        native = CT.encode_alignments(rows, min_mq, dcov)
        python = CT.encode_alignments(rows, min_mq, dcov, encoder="python")
        assert native.n_reads > 50
        for field in CT.Alignments.__slots__:
            assert np.array_equal(getattr(native, field), getattr(python, field)), field


def test_native_encoder_rejects_malformed_rows():
    with pytest.raises(ValueError):
        CT.encode_alignments(["r\t0\tc\tnot_a_number\t60\t1M\t*\t0\t0\tA\t*"])
    with pytest.raises(ValueError):
        CT.encode_alignments(["r\t0\tc\t5\t60\t1M"])


def test_device_tensors_stand_in():
    x = np.arange(6 * 1056, dtype=np.int16).reshape(6, 33, 8, 4)
    block = CT.TensorBlock(None, "c", np.array([100, 150, 200, 250, 300, 350]), "ACGTN*" * 100, np.array([83, 133, 183, 233, 283, 333]),
                           np.zeros(6, np.int32), np.arange(6), x, True)
    assert block.sequences[0] == ("ACGTN*" * 100)[83:116] and len(block.sequences[0]) == 33
    # centre bases (window start + 16): indices 99, 149, ... of the text; '*' is not an IUPAC code (clair/utils.py:90)
    text = "ACGTN*" * 100
    keep = block.callable_sites()
    assert keep.tolist() == [i for i, s in enumerate([83, 133, 183, 233, 283, 333]) if text[s + 16] != "*"]
    assert len(keep) == 4
    batches = list(CT.device_tensor_generator_from(block, 3))
    assert [len(X) for X, _ in batches] == [3, 1] and batches[0][0].shape == (3, 33, 8, 4)
    X, infos = batches[0]
    assert infos[0] == ["c", str(int(block.positions[keep[0]])), block.sequences[keep[0]]]
    assert np.array_equal(np.stack(list(X)), x[keep[:3]]) and np.array_equal(X[1], x[keep[1]])
    no_host = CT.DeviceTensors(CT.TensorBlock(None, "c", block.positions, text, block._start, block.depth, block.rows, None, True), keep)
    with pytest.raises(ValueError):
        list(no_host)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="runs the reference's own batch_output; only where /root/reference exists")
def test_reference_batch_output_accepts_device_tensors():
    """The reference's unmodified output stage (clair/call_var.py:1199-1236 -> output_with -> output_from) prints the same VCF
    rows from a DeviceTensors batch as from the numpy batch tensor_generator_from would have handed it."""
    from oracle import gen_golden_decision as GD
    from clair_b200 import synth
    cv = GD.import_reference_call_var()
    rng = np.random.default_rng(5)
    n = 120
    x = synth.synthetic_counts(n, seed=21)
    x[..., 1:] -= x[..., 0:1]
    text = "".join(rng.choice(list("ACGT"), size=50 * n + 100))
    positions = np.arange(n, dtype=np.int64) * 50 + 40
    block = CT.TensorBlock(None, "chr7", positions, text, positions - 17, np.zeros(n, np.int32), np.arange(n), x, True)
    (X, infos), = list(CT.device_tensor_generator_from(block, 1000))
    P = np.stack([np.concatenate([GD.softmax(rng.normal(0, 3.0, k)) for k in (21, 3, 33, 33)]) for _ in range(n)])
    batch_Y = [P[:, 0:21], P[:, 21:24], P[:, 24:57], P[:, 57:90]]
    config = cv.OutputConfig(is_show_reference=True, is_debug=False, is_haploid_precision_mode_enabled=False,
                             is_haploid_sensitive_mode_enabled=False, is_output_for_ensemble=False, quality_score_for_pass=None)

    def run(batch_x):
        lines, rec = [], GD.Recorder()
        util = cv.OutputUtilities(print_debug_message=lambda *a: lines.append(("debug", a[0], a[1], a[-1])),
                                  insertion_bases_using=rec.insertion_bases_using, deletion_bases_using=rec.deletion_bases_using,
                                  insertion_bases_using_pysam_using=rec.insertion_bases_using_pysam_using,
                                  output=lines.append, output_header=lambda: None, close_opened_files=lambda: None)
        cv.batch_output((batch_x, infos), batch_Y, config, util)
        return lines

    from_numpy = run(x[block.callable_sites()].astype(np.float32))
    from_device_batch = run(X)
    assert len(from_numpy) > n // 2 and from_device_batch == from_numpy


# ---- the reference's command line (host plumbing; the device call is replaced by the oracle in THIS test only) ---------
def _fake_children(case):
    contig = case["contig"]

    def popen(args, **kw):
        if "faidx" in args:
            region = args[-1]
            lo, hi = 1, len(contig)
            if ":" in region:
                lo, hi = (int(v) for v in region.split(":")[1].split("-"))
            seq = contig[lo - 1:hi]
            return G._Proc([">%s\n" % region] + [seq[i:i + 60].lower() + "\n" for i in range(0, len(seq), 60)])
        if "view" in args:
            assert "-F" in args and str(CT.param.SAMTOOLS_VIEW_FILTER_FLAG) in args
            return G._Proc([l + "\n" for l in case["sam"]])
        if args[0] == "gzip":
            return G._Proc([l + "\n" for l in case["candidates"]])
        raise AssertionError("unexpected child process %r" % (args,))
    return popen


@pytest.mark.parametrize("name", ["whole", "region", "min_mq_cov", "no_left_edge", "contig_tail"])
def test_output_aln_tensor_plumbing(monkeypatch, tmp_path, name):
    """OutputAlnTensor(args): children, region strings, filters and row formatting (CreateTensor.py:179-394) with the counting
    done by the oracle instead of the device - what is left is exactly the host code around clairb_create_tensors."""
    import types
    case = next(c for c in CASES if c["name"] == name)
    a = case["args"]

    def oracle_create_tensors(model, alignments, candidates, reference_sequence, start0=0, ctg_name="chr", min_coverage=0,
                              consider_left_edge=True, ctg_start=None, ctg_end=None, subtract=False, fetch=True):
        rows = O.create_tensors(case["sam"], sorted(set(candidates)), reference_sequence, start0, ctg_name, a["minMQ"], a["dcov"],
                                min_coverage, consider_left_edge, ctg_start, ctg_end)
        positions = np.array([r[1] for r in rows], np.int64)
        x = np.stack([r[3] for r in rows]).astype(np.int16) if rows else np.zeros((0, 33, 8, 4), np.int16)
        return CT.TensorBlock(model, ctg_name, positions, reference_sequence, positions - start0 - 17, x[:, 16, :, 0].sum(1),
                              np.arange(len(rows)), x, False)

    monkeypatch.setattr(CT, "create_tensors", oracle_create_tensors)
    monkeypatch.setattr(CT.param, "expandReferenceRegion", a["expandReferenceRegion"])
    args = types.SimpleNamespace(samtools="samtools", tensor_fn=str(tmp_path / "t.gz"), bam_fn="x.bam", ref_fn="x.fa", can_fn="x.can",
                                 dcov=a["dcov"], stop_consider_left_edge=a["stop_consider_left_edge"], minCoverage=a["minCoverage"],
                                 minMQ=a["minMQ"], ctgName=a["ctgName"], ctgStart=a["ctgStart"], ctgEnd=a["ctgEnd"])
    out = []
    CT.OutputAlnTensor(args, model=object(), popen=_fake_children(case), out=out)
    assert out == case["expected"]
    CT.OutputAlnTensor(args, model=object(), popen=_fake_children(case))                 # --tensor_fn FILE: gzip text
    with gzip.open(args.tensor_fn, "rt") as f:
        assert f.read().splitlines() == case["expected"]


# ---- GPU -------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def model():
    from clair_b200 import weights as W
    from clair_b200.model import Clair
    m = Clair(max_sites=4096, batch_sites=1000)
    m.set_weights(W.random_weights(seed=1234))
    yield m
    m.close()


def device_block(model, case, **kw):
    a = case["args"]
    seq, start0 = reference_window(case)
    aln = CT.encode_alignments(case["sam"], a["minMQ"], a["dcov"])
    return CT.create_tensors(model, aln, candidates_of(case), seq, start0, a["ctgName"], a["minCoverage"],
                             not a["stop_consider_left_edge"], a["ctgStart"], a["ctgEnd"], **kw)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=CASE_IDS)
def test_device_rows_equal_reference_rows(model, case):
    assert device_block(model, case).text_rows() == case["expected"]


@pytest.mark.gpu
def test_device_random_regions_equal_oracle(model):
    rng = np.random.default_rng(78)
    for k in range(4):
        case = G.make_case(rng, "big%d" % k, 6000, 1500, 700, style=["mixed", "dense", "leading", "short"][k],
                           stop_left=(k == 2), dup_rate=0.2, dcov=5)
        want = oracle_rows(case)
        block = device_block(model, case)
        assert block.positions.tolist() == [r[1] for r in want] and block.sequences == [r[2] for r in want]
        assert np.array_equal(block.x, np.stack([r[3] for r in want]).astype(np.int16))
        # the subtracted form is what tensor_generator_from hands the network (clair/utils.py:96-98)
        sub = device_block(model, case, subtract=True)
        expect = block.x.copy()
        expect[..., 1:] -= expect[..., 0:1]
        assert np.array_equal(sub.x, expect)


@pytest.mark.gpu
def test_device_born_tensors_through_forward(model):
    rng = np.random.default_rng(79)
    case = G.make_case(rng, "fwd", 5000, 2500, 600)
    block = device_block(model, case, subtract=True)
    keep = block.callable_sites()
    assert 0 < keep.shape[0] <= len(block)
    got = block.predict(keep)
    want = model.predict_packed(block.x[keep])
    assert np.array_equal(got, want)
    batches = list(CT.created_tensor_generator_from(block, 250))
    assert sum(len(info) for _, info in batches) == keep.shape[0]
    assert np.array_equal(np.concatenate([p for p, _ in batches]), want)
    assert batches[0][1][0] == [block.ctg_name, str(int(block.positions[keep[0]])), block.sequences[keep[0]]]


@pytest.mark.gpu
def test_batch_loop_over_device_born_tensors(model):
    """call_var's loop (clair/call_var.py:1312-1367) fed by device_tensor_generator_from: same batches, same order, same
    probabilities as the text route (tensor rows -> tensor_generator_from -> predict)."""
    from clair_b200 import call_var

    class Utilities(object):
        calls = []

        def output_header(self):
            self.calls.append("header")

        def close_opened_files(self):
            self.calls.append("close")

    rng = np.random.default_rng(80)
    case = G.make_case(rng, "loop", 5000, 2500, 600)
    block = device_block(model, case, subtract=True)
    keep = block.callable_sites()
    seen = []

    def output_stage(batch, prediction, config, utilities):
        X, infos = batch
        assert len(X) == len(infos) == prediction[0].shape[0]
        seen.append((infos, [p.copy() for p in prediction]))

    util = Utilities()
    call_var.call_variants_from_alignments(block, model, None, util, output_stage)
    assert util.calls == ["header", "close"]
    assert [i[1] for infos, _ in seen for i in infos] == [str(int(p)) for p in block.positions[keep]]
    want = model.predict(block.x[keep])
    for k in range(4):
        assert np.array_equal(np.concatenate([pred[k] for _, pred in seen]), want[k])
    assert [len(infos) for infos, _ in seen] == [1000] * (keep.shape[0] // 1000) + ([keep.shape[0] % 1000] if keep.shape[0] % 1000 else [])


@pytest.mark.gpu
def test_device_argument_errors(model):
    case = CASES[0]
    seq, start0 = reference_window(case)
    aln = CT.encode_alignments(case["sam"])
    aln.read_pos = aln.read_pos[::-1].copy()
    with pytest.raises(ValueError):
        CT.create_tensors(model, aln, candidates_of(case), seq, start0)
    # malformed ops are found on the device (validate_ops) before the counting kernel reads SEQ through them
    aln = CT.encode_alignments(case["sam"])
    aligned = np.flatnonzero((aln.op_len & 3) != 2)                      # ops that read SEQ (M = 0, I = 1; D = 2 does not)
    bad = CT.encode_alignments(case["sam"])
    bad.op_qry = aln.op_qry.copy()
    bad.op_qry[aligned[-1]] = aln.seq.size                              # starts at the end of SEQ
    with pytest.raises(ValueError, match="past the end of its SEQ"):
        CT.create_tensors(model, bad, candidates_of(case), seq, start0)
    bad = CT.encode_alignments(case["sam"])
    bad.op_len = aln.op_len.copy()
    bad.op_len[0] = aln.op_len[0] | 3                                   # the encoder only writes codes 0..2
    with pytest.raises(ValueError, match="unknown code"):
        CT.create_tensors(model, bad, candidates_of(case), seq, start0)
    # and the handle still works afterwards
    assert device_block(model, case).text_rows() == case["expected"]
