"""Golden VCF rows produced by the REFERENCE's own `output_with` (clair/call_var.py:1002-1197).  TEST INFRASTRUCTURE.

Run from the repo root (only where /root/reference exists):  python oracle/gen_golden_output.py
Imports /root/reference/clair/call_var.py (pysam / tensorflow / blosc / intervaltree stubbed as in gen_golden_decision.py)
and calls its unmodified `output_with` for every case under several output configurations; what it hands to
`output_utilities.output` / `print_debug_message` is stored in tests/golden/output_rows.json.gz next to the inputs.

Scalar semantics: the reference pins numpy 1.18 (README.md:127), where a numpy.float32 scalar combined with a Python
float promotes to float64 - `1.0 - p` and `p + 1e-300` in quality_score_from (:581-583) rely on it (under numpy >= 2 the
same lines stay in float32, and `log` raises for p == 1).  This container has numpy 2, so the probability vectors are
handed to the reference as an ndarray subclass whose scalar items behave like numpy-1.18 float32 scalars: float32 product
with each other, float64 with Python floats.  Array arithmetic is untouched (identical in both numpy generations), the
tensors go in as float64 (integer counts: exact in every float type).  The reference's source is not modified.
"""
import gzip
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

from oracle import gen_golden_decision as GD  # noqa: E402


class L32(float):
    """A float32 scalar as numpy 1.18 treats it: float32 * float32 -> float32, anything with a Python float -> float64."""

    def __mul__(self, other):
        if isinstance(other, np.ndarray):
            return NotImplemented
        if isinstance(other, (L32, np.float32)):
            return L32(float(np.float32(self) * np.float32(other)))
        return float(self) * other

    __rmul__ = __mul__


class LegacyArray(np.ndarray):
    def __getitem__(self, index):
        item = super().__getitem__(index)
        return L32(float(item)) if isinstance(item, np.floating) else item


def legacy(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(LegacyArray)


class Stingy(GD.Recorder):
    """indel-base helpers that come back empty for some lengths: output_from's loop then moves on to its next candidate"""

    def insertion_bases_using(self, tensor_input, variant_length, contig, position):
        return ("", 0) if variant_length % 3 == 0 else GD.Recorder.insertion_bases_using(self, tensor_input, variant_length, contig, position)

    def deletion_bases_using(self, tensor_input, variant_length, contig, position, reference_sequence):
        return ("", 0) if variant_length % 4 == 0 else GD.Recorder.deletion_bases_using(self, tensor_input, variant_length, contig, position, reference_sequence)


CONFIGS = [
    dict(is_show_reference=True, is_debug=False, is_haploid_precision_mode_enabled=False, is_haploid_sensitive_mode_enabled=False,
         is_output_for_ensemble=False, quality_score_for_pass=None),
    dict(is_show_reference=False, is_debug=False, is_haploid_precision_mode_enabled=False, is_haploid_sensitive_mode_enabled=False,
         is_output_for_ensemble=False, quality_score_for_pass=180),
    dict(is_show_reference=True, is_debug=False, is_haploid_precision_mode_enabled=True, is_haploid_sensitive_mode_enabled=False,
         is_output_for_ensemble=False, quality_score_for_pass=748),
    dict(is_show_reference=False, is_debug=False, is_haploid_precision_mode_enabled=False, is_haploid_sensitive_mode_enabled=True,
         is_output_for_ensemble=False, quality_score_for_pass=None),
]


def make_inputs(rng):
    """Probability vectors, centre rows of the tensors and the info triples of the cases."""
    from clair_b200 import synth
    P, _ = GD.make_cases(rng)
    # call sets look different from random heads: add confident reference / SNP calls (p close to 1 is where the float32
    # product matters most for the quality score) and a few p == 1 / p == 0 heads
    extra = []
    for sharp in (9.0, 14.0, 20.0, 40.0):
        for _ in range(120):
            z = [rng.normal(0, 1.0, k) for k in (21, 3, 33, 33)]
            z[0][rng.integers(0, 10)] += sharp
            z[1][rng.integers(0, 3)] += sharp
            z[2][16] += sharp
            z[3][16] += sharp
            extra.append(np.concatenate([GD.softmax(v) for v in z]))
    P = np.concatenate([P, np.stack(extra).astype(np.float32)])
    n = len(P)
    X = synth.synthetic_tensors(n, seed=4242)                      # channel-subtracted, as the generator yields them
    X[5] = 0                                                       # zero read depth
    X[6, 16] = 0
    bases = rng.integers(0, 4, n)
    infos = []
    for i in range(n):
        seq = list("".join(rng.choice(list("ACGT"), size=33)))
        seq[16] = "ACGT"[bases[i]]
        if i % 97 == 11:
            seq[16] = "N"                                          # not a basic base: no row (:1012-1013)
        if i % 97 == 12:
            seq[16] = "U"
        infos.append(["chr%d" % (1 + i % 3), str(1000 + 37 * i), "".join(seq)])
    return P, X, infos


def main():
    if not os.path.isdir(GD.REFERENCE):
        print("no /root/reference here: output fixtures not regenerated")
        return
    cv = GD.import_reference_call_var()
    rng = np.random.default_rng(20240618)
    P, X, infos = make_inputs(rng)
    n = len(P)
    runs = []
    for ci, cfg in enumerate(CONFIGS):
        for helper_name, helper in (("recorder", GD.Recorder), ("stingy", Stingy)):
            if helper_name == "stingy" and ci not in (0, 1):
                continue
            config = cv.OutputConfig(**cfg)
            rows = []
            for i in range(n):
                out, rec = [], helper()
                util = cv.OutputUtilities(print_debug_message=lambda *a: out.append(["debug", a[-1]]),
                                          insertion_bases_using=rec.insertion_bases_using, deletion_bases_using=rec.deletion_bases_using,
                                          insertion_bases_using_pysam_using=rec.insertion_bases_using_pysam_using,
                                          output=lambda s: out.append(["row", s]), output_header=lambda: None,
                                          close_opened_files=lambda: None)
                p = P[i]
                cv.output_with(X[i].astype(np.float64), infos[i], legacy(p[0:21]), legacy(p[21:24]), legacy(p[24:57]),
                               legacy(p[57:90]), config, util)
                assert len(out) <= 1
                rows.append(out[0] if out else None)
            runs.append({"config": cfg, "helpers": helper_name, "rows": rows})
            print("config %d helpers %-8s: %d rows, %d debug messages, %d silent" % (
                ci, helper_name, sum(1 for r in rows if r and r[0] == "row"), sum(1 for r in rows if r and r[0] == "debug"),
                sum(1 for r in rows if r is None)))
    blob = {"probs_f32_hex": P.astype("<f4").tobytes().hex(), "x_rows_16_17_i16_hex": X[:, 16:18].astype("<i2").tobytes().hex(),
            "n": n, "infos": infos, "runs": runs}
    with gzip.open(os.path.join(GOLD, "output_rows.json.gz"), "wt") as f:
        json.dump(blob, f)
    print("output_rows.json.gz: %d sites x %d runs" % (n, len(runs)))


if __name__ == "__main__":
    main()
