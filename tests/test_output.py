"""The batched VCF output stage (clair_b200/output.py) against rows printed by the REFERENCE's own
output_with (tests/golden/output_rows.json.gz, made by oracle/gen_golden_output.py from /root/reference).

CPU: the oracle restatements (oracle/output_oracle.py, the quality / support derivation of oracle/decision_oracle.py) and
the product stage fed with oracle-made decision records.  GPU (-m gpu): the same stage fed with the device's records.
"""
import ctypes
import gzip
import json
import os
import types

import numpy as np
import pytest

from clair_b200 import _lib, decision, output
from oracle import decision_oracle as DO
from oracle import output_oracle as OO

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "output_rows.json.gz")


@pytest.fixture(scope="module")
def gold():
    with gzip.open(GOLDEN, "rt") as f:
        g = json.load(f)
    n = g["n"]
    g["P"] = np.frombuffer(bytes.fromhex(g["probs_f32_hex"]), "<f4").reshape(n, 90).copy()
    rows = np.frombuffer(bytes.fromhex(g["x_rows_16_17_i16_hex"]), "<i2").reshape(n, 2, 8, 4)
    X = np.zeros((n, 33, 8, 4), np.float32)
    X[:, 16:18] = rows
    g["X"] = X
    g["ref_codes"] = decision.ref_base_codes(g["infos"])
    # exact ties between categories: the reference's flags tuple then has several True entries and its elif chains may
    # follow another one than the category REF / ALT were built from (the documented difference of decision.FirstChoice)
    g["tied"] = np.array([DO.categories_holding_the_maximum(g["P"][i], g["ref_codes"][i]) > 1 for i in range(n)])
    return g


@pytest.fixture(scope="module")
def oracle_records(gold):
    dec, maxp, depth, quality, support = DO.decide_full(gold["P"], gold["ref_codes"], gold["X"])
    return decision.Decision(dec[:, 0], dec[:, 1], dec[:, 2], dec[:, 3], maxp, depth, quality, support)


class Helpers:
    """The indel-base helpers the golden runs used (oracle/gen_golden_output.py: GD.Recorder / Stingy), restated here so the
    GPU box needs nothing from oracle/gen_*."""

    def __init__(self, stingy):
        self.stingy = stingy

    def insertion_bases_using(self, tensor_input, variant_length, contig, position):
        if self.stingy and variant_length % 3 == 0:
            return "", 0
        return "A" * variant_length, variant_length

    def deletion_bases_using(self, tensor_input, variant_length, contig, position, reference_sequence):
        if self.stingy and variant_length % 4 == 0:
            return "", 0
        return ("CGTA" * 5)[:variant_length], variant_length

    def insertion_bases_using_pysam_using(self, contig, position, minimum_insertion_length, maximum_insertion_length,
                                          insertion_bases_to_ignore):
        return "C" * minimum_insertion_length


def utilities(stingy, sink):
    h = Helpers(stingy)
    return types.SimpleNamespace(insertion_bases_using=h.insertion_bases_using, deletion_bases_using=h.deletion_bases_using,
                                 insertion_bases_using_pysam_using=h.insertion_bases_using_pysam_using,
                                 output=lambda s: sink.extend(["row", r] for r in s.split("\n")),
                                 print_debug_message=lambda *a: sink.append(["debug", a[-1]]))


def run_stage(gold, dec, run, batch=257):
    """BatchOutput over the golden sites in batches -> per site what came out (None / ["row", s] / ["debug", msg] / "fallback")."""
    cfg = types.SimpleNamespace(**run["config"])
    n = gold["n"]
    P, X, infos = gold["P"], gold["X"], gold["infos"]
    got = [None] * n
    for lo in range(0, n, batch):
        hi = min(n, lo + batch)
        sink = []

        def fallback(x, info, *rest, _lo=lo):
            sink.append(["fallback", info[1]])

        stage = output.BatchOutput(cfg, utilities(run["helpers"] == "stingy", sink), fallback=fallback)
        sub = decision.Decision(*[np.asarray(f)[lo:hi] for f in dec])
        stage((X[lo:hi], infos[lo:hi]), [P[lo:hi, 0:21], P[lo:hi, 21:24], P[lo:hi, 24:57], P[lo:hi, 57:90]], sub)
        by_pos = {info[1]: k for k, info in enumerate(infos[lo:hi], lo)}
        last = -1
        for kind, payload in sink:
            if kind == "row":
                k = by_pos[payload.split("\t")[1]]
            elif kind == "fallback":
                k = by_pos[payload]
            else:                                                    # debug messages carry no position: zero-depth sites
                k = next(j for j in range(max(last + 1, lo), hi) if gold["X"][j, 16].any() == 0 and got[j] is None and
                         infos[j][2][16] in "ACGTU")
            assert k > last, "rows must leave in site order"
            last = k
            got[k] = "fallback" if kind == "fallback" else [kind, payload]
    return got


def check_against_golden(gold, dec):
    total = fell = 0
    for run in gold["runs"]:
        got = run_stage(gold, dec, run)
        for i, (g, want) in enumerate(zip(got, run["rows"])):
            if g == "fallback":
                fell += 1                                            # the reference's own output_with prints these
                continue
            if gold["tied"][i]:
                continue
            assert g == want, "site %d (%s) under %s / %s: got %r, reference printed %r" % (
                i, gold["infos"][i], run["config"], run["helpers"], g, want)
            total += 1
    return total, fell


def test_oracle_quality_and_support_reproduce_the_reference_rows(gold, oracle_records):
    # oracle/output_oracle.output_row fed with the first choice (decision oracle + FirstChoice string assembly) prints the
    # reference's rows; the category-derived quality / support of oracle/decision_oracle.py equal the string-derived ones
    run = gold["runs"][0]
    assert run["helpers"] == "recorder" and run["config"]["is_show_reference"]
    util = utilities(False, [])
    checked = 0
    for i in range(gold["n"]):
        info, want = gold["infos"][i], run["rows"][i]
        if info[2][16] not in "ACGT" or not gold["X"][i, 16].any() or gold["tied"][i]:
            continue
        c, l1, l2, aux = (int(np.asarray(f)[i]) for f in (oracle_records.category, oracle_records.len1, oracle_records.len2, oracle_records.aux))
        answer = decision.FirstChoice._first_choice(gold["X"][i], info[2], info[0], int(info[1]), 16, util, c, l1, l2, aux)
        if answer is None:
            continue
        flags, (ref, alt) = answer
        p = gold["P"][i]
        row = OO.output_row(gold["X"][i], info, p[0:21], p[21:24], flags, ref, alt)
        assert (["row", row] if row is not None else None) == want
        if row is not None:
            fields = row.split("\t")
            assert int(fields[5]) == oracle_records.quality[i]
            assert OO.supported_reads(gold["X"][i], flags, ref, alt) == float(oracle_records.supported_reads[i])
            checked += 1
    assert checked > 1000


def test_batch_output_prints_the_reference_rows_from_oracle_records(gold, oracle_records):
    total, fell = check_against_golden(gold, oracle_records)
    assert total > 6500 and 0 < fell < total // 10
    cats = set(np.asarray(oracle_records.category).tolist())
    assert cats == set(range(10))


def test_batch_output_without_fallback_says_which_site_needs_it(gold, oracle_records):
    run = next(r for r in gold["runs"] if r["helpers"] == "stingy")
    stage = output.BatchOutput(types.SimpleNamespace(**run["config"]), utilities(True, []))
    P = gold["P"]
    with pytest.raises(output.NeedsFallback, match="first choice does not stand"):
        stage((gold["X"], gold["infos"]), [P[:, 0:21], P[:, 21:24], P[:, 24:57], P[:, 57:90]], oracle_records)
    with pytest.raises(ValueError):
        output.BatchOutput(types.SimpleNamespace(is_output_for_ensemble=True), None)


def test_native_vcf_row_formatter_equals_python_formatting():
    lib = _lib.load()
    rng = np.random.default_rng(3)
    n = 500
    names = [b"chr%d" % rng.integers(1, 23) if i % 7 else b"a_rather_long_contig_name.1" for i in range(n)]
    off = np.zeros(n + 1, np.int32)
    np.cumsum([len(s) for s in names], out=off[1:])
    pos = rng.integers(1, 3_000_000_000, n).astype(np.int64)
    ref = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)].copy()
    alt = np.zeros((n, 4), np.uint8)
    alt[:, 0] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)]
    multi = rng.random(n) < 0.3
    alt[multi, 1] = ord(",")
    alt[multi, 2] = ord("T")
    quality = rng.integers(0, 9_200_000, n).astype(np.int32)
    filt = rng.integers(0, 3, n).astype(np.uint8)
    gt = rng.integers(0, 6, n).astype(np.uint8)
    depth = rng.integers(1, 5000, n).astype(np.int32)
    af = np.concatenate([rng.random(n - 6), [0.0, 1.0, 0.00005, 0.99995, 0.12345, 0.5]]).astype(np.float64)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    args = (n, b"".join(names), p(off), p(pos), p(ref), p(alt), p(quality), p(filt), p(gt), p(depth), p(af))
    need = ctypes.c_int64()
    assert lib.clairb_format_vcf_rows(*args, None, 0, ctypes.byref(need), None) == 0
    out = ctypes.create_string_buffer(need.value + 1)
    row_end = np.empty(n, np.int64)
    assert lib.clairb_format_vcf_rows(*args, out, need.value, ctypes.byref(need), p(row_end)) == _lib.EINVAL      # one byte short
    assert lib.clairb_format_vcf_rows(*args, out, need.value + 1, ctypes.byref(need), p(row_end)) == 0
    rows = out.raw[:need.value].decode().split("\n")
    F, G = [".", "PASS", "LowQual"], ["0/0", "1/1", "0/1", "1/2", "0", "1"]
    for i in range(n):
        a = bytes(alt[i]).split(b"\0")[0].decode()
        want = "%s\t%d\t.\t%s\t%s\t%d\t%s\t%s\tGT:GQ:DP:AF\t%s:%d:%d:%.4f" % (
            names[i].decode(), pos[i], chr(ref[i]), a, quality[i], F[filt[i]], ".", G[gt[i]], quality[i], depth[i], af[i])
        assert rows[i] == want
        assert out.raw[:row_end[i]].decode().split("\n")[-1] == want
    bad = gt.copy()
    bad[3] = 9
    assert lib.clairb_format_vcf_rows(n, b"".join(names), p(off), p(pos), p(ref), p(alt), p(quality), p(filt), p(bad), p(depth), p(af),
                                      None, 0, ctypes.byref(need), None) == _lib.EINVAL


@pytest.mark.gpu
def test_device_records_print_the_reference_rows(gold, oracle_records, gpu_model):
    # words 6 / 7 of the device's decision records (quality score, supporting reads) against the oracle derivation, and the
    # whole stage on the device's records against the rows the reference printed
    dev = gpu_model.decide(gold["P"], gold["ref_codes"], gold["X"])
    for f in ("category", "len1", "len2", "aux", "max_probability", "read_depth", "quality", "supported_reads"):
        np.testing.assert_array_equal(getattr(dev, f), getattr(oracle_records, f), err_msg=f)
    total, fell = check_against_golden(gold, dev)
    assert total > 6500
    dev16 = gpu_model.decide(gold["P"], gold["ref_codes"], gold["X"].astype(np.int16))     # int16 transport of the same counts
    np.testing.assert_array_equal(dev16.quality, dev.quality)
    np.testing.assert_array_equal(dev16.supported_reads, dev.supported_reads)


@pytest.mark.gpu
def test_batch_output_behind_the_batch_loop_on_the_device(gpu_model, weights1234):
    # run_batches(with_decision=True) -> BatchOutput: rows of every batch, in order, equal to the per-site oracle rows
    from clair_b200 import call_var, synth
    sizes = [1000, 1000, 400]
    rng = np.random.default_rng(17)
    items = []
    for k, s in enumerate(sizes):
        X = synth.synthetic_tensors(s, seed=700 + k)
        infos = [["chr7", str(10_000 * k + 3 * j + 1), "".join(rng.choice(list("ACGT"), size=33))] for j in range(s)]
        items.append((X, infos))
    sink = []
    cfg = types.SimpleNamespace(is_show_reference=True, is_debug=False, is_haploid_precision_mode_enabled=False,
                                is_haploid_sensitive_mode_enabled=False, is_output_for_ensemble=False, quality_score_for_pass=200)
    stage = output.BatchOutput(cfg, utilities(False, sink), fallback=lambda x, info, *rest: sink.append(["fallback", info[1]]))
    call_var.run_batches(gpu_model, iter(items), stage, with_decision=True, in_flight=4)
    assert stage.fast_rows + stage.slow_rows + stage.fallback_sites > 2000
    positions = [int(r[1].split("\t")[1]) if r[0] == "row" else int(r[1]) for r in sink if r[0] != "debug"]
    assert positions == sorted(positions)
    # spot-check rows against the per-site oracle on the device's probabilities
    util = utilities(False, [])
    X, infos = items[1]
    pred, dec = gpu_model.predict_and_decide(X, decision.ref_base_codes(infos))
    rows = {r[1].split("\t")[1]: r[1] for r in sink if r[0] == "row"}
    checked = 0
    for i in range(0, 1000, 7):
        answer = decision.FirstChoice._first_choice(X[i], infos[i][2], "chr7", int(infos[i][1]), 16, util, int(dec.category[i]),
                                                    int(dec.len1[i]), int(dec.len2[i]), int(dec.aux[i]))
        if answer is None:
            continue
        flags, (ref, alt) = answer
        want = OO.output_row(X[i], infos[i], pred[0][i], pred[1][i], flags, ref, alt, quality_score_for_pass=200)
        assert rows.get(infos[i][1]) == want
        checked += want is not None
    assert checked > 100
