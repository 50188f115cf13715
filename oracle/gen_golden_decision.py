"""Golden vectors for the first-choice variant decision, produced by the REFERENCE's own code.  TEST INFRASTRUCTURE.

Run from the repo root (only where /root/reference exists):  python oracle/gen_golden_decision.py
Imports /root/reference/clair/call_var.py with its absent third-party imports (pysam, tensorflow via clair.model,
blosc, intervaltree) stubbed, then for every case calls
    call_var.output_from(...)                      -> the flags tuple (category) and, through recording stubs of the
                                                      indel-base helpers, the variant lengths the reference chose
    call_var.possible_outcome_probabilites_from    -> the hetero base of ACGT+Ins / ACGT+Del outcomes
    call_var.homo_SNP_bases_from / hetero_...      -> the SNP label
and stores them in tests/golden/decision_cases.npz next to the inputs.
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"


def import_reference_call_var():
    sys.path.insert(0, REFERENCE)
    for name in ("blosc", "intervaltree", "pysam"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.set_nthreads = lambda n: None
            m.IntervalTree = object
            sys.modules[name] = m
    if "clair.model" not in sys.modules:
        import clair                                   # the reference package
        stub = types.ModuleType("clair.model")
        stub.Clair = object
        sys.modules["clair.model"] = stub
        clair.model = stub
    import clair.call_var as cv
    return cv


def softmax(z):
    z = z - z.max()
    e = np.exp(z)
    return (e / e.sum()).astype(np.float32)


def make_cases(rng):
    """probability vectors [n,90] float32 (four softmax heads side by side) and reference bases 0..3"""
    cases = []
    # 1) softmax of random logits at several sharpnesses: exercises every category
    for sharp in (0.5, 1.5, 3.0, 6.0, 12.0):
        for _ in range(90):
            p = np.concatenate([softmax(rng.normal(0, sharp, k)) for k in (21, 3, 33, 33)])
            cases.append(p)
    # 2) heads steered towards one category each (peaked gt21 label x genotype x lengths)
    for gt in range(21):
        for geno in range(3):
            for _ in range(3):
                z21 = rng.normal(0, 1, 21); z21[gt] += 6
                z3 = rng.normal(0, 1, 3); z3[geno] += 4
                l1 = rng.normal(0, 1, 33); l1[16 if gt < 10 else rng.integers(0, 33)] += 5
                l2 = rng.normal(0, 1, 33); l2[16 if gt < 10 else rng.integers(0, 33)] += 5
                cases.append(np.concatenate([softmax(z21), softmax(z3), softmax(l1), softmax(l2)]))
    # 3) ties: uniform heads, exact one-hot heads, duplicated maxima
    cases.append(np.concatenate([np.full(k, 1.0 / k, np.float32) for k in (21, 3, 33, 33)]))
    for gt in range(21):
        p = np.zeros(90, np.float32)
        p[gt] = 1.0
        p[21 + gt % 3] = 1.0
        p[24 + (gt * 5) % 33] = 1.0
        p[57 + (gt * 7) % 33] = 1.0
        cases.append(p)
    for _ in range(40):
        p = np.concatenate([softmax(rng.normal(0, 2.0, k)) for k in (21, 3, 33, 33)])
        i, j = rng.integers(24, 57, 2)
        p[j] = p[i]                                    # equal length probabilities -> equal products
        i, j = rng.integers(0, 21, 2)
        p[j] = p[i]
        cases.append(p)
    # 4) probabilities restricted to a few powers of two: thousands of exactly equal products per site, the winner is
    #    decided by the reference's tie order alone
    levels = np.array([0.0, 0.125, 0.25, 0.5], np.float32)
    for k in range(150):
        p = levels[rng.integers(0, 4, size=90)]
        if k < 15:
            p[24:90] = 0.25
        elif k < 30:
            p[0:21] = 0.5
        cases.append(p)
    cases.append(np.zeros(90, np.float32))             # everything zero: the reference answers "reference"
    P = np.stack(cases).astype(np.float32)
    ref_bases = rng.integers(0, 4, len(P)).astype(np.uint8)
    return P, ref_bases


class Recorder(object):
    """Stands in for OutputUtilities' indel-base helpers: records the variant lengths it is asked for and returns
    non-empty, mutually different bases so that output_from accepts its first choice (no retry)."""

    def __init__(self):
        self.calls = []

    def insertion_bases_using(self, tensor_input, variant_length, contig, position):
        self.calls.append(("ins", variant_length))
        return "A" * variant_length, variant_length

    def deletion_bases_using(self, tensor_input, variant_length, contig, position, reference_sequence):
        self.calls.append(("del", variant_length))
        return ("CGTA" * 5)[:variant_length], variant_length

    def insertion_bases_using_pysam_using(self, contig, position, minimum_insertion_length, maximum_insertion_length,
                                           insertion_bases_to_ignore):
        self.calls.append(("ins_pysam", minimum_insertion_length))
        return "C" * minimum_insertion_length


def main():
    if not os.path.isdir(REFERENCE):
        print("no /root/reference here: decision fixtures not regenerated")
        return
    cv = import_reference_call_var()
    rng = np.random.default_rng(20240611)
    P, ref_bases = make_cases(rng)
    n = len(P)
    X = rng.integers(-30, 60, size=(n, 33, 8, 4)).astype(np.float32)
    dec = np.zeros((n, 4), np.int32)
    maxp = np.zeros(n, np.float32)
    depth = np.zeros(n, np.float32)
    for k in range(n):
        gt21, geno, vl1, vl2 = P[k, 0:21], P[k, 21:24], P[k, 24:57], P[k, 57:90]
        base = "ACGT"[ref_bases[k]]
        seq = "T" * 16 + base + "G" * 16
        rec = Recorder()
        util = types.SimpleNamespace(insertion_bases_using=rec.insertion_bases_using,
                                     deletion_bases_using=rec.deletion_bases_using,
                                     insertion_bases_using_pysam_using=rec.insertion_bases_using_pysam_using)
        flags, (ref_out, alt_out) = cv.output_from(X[k], seq, "chr1", 1000 + k, 16, gt21, geno, vl1, vl2, None, util)
        cat = list(flags).index(True)
        lists = cv.possible_outcome_probabilites_from(gt21, geno, vl1, vl2, reference_base=base)
        maximum = max([lists[0], max(lists[1]), max(lists[2]), max(lists[4]), max(lists[11]), max(lists[9]),
                       max(lists[6]), max(lists[16]), max(lists[13]), max(lists[18])])
        l1 = l2 = aux = 0
        calls = dict((kind, v) for kind, v in reversed(rec.calls))     # first call of each kind wins
        if cat == 0:
            aux = cv.gt21_enum_from_label(base + base)
        elif cat == 1:
            b1, b2 = cv.homo_SNP_bases_from(gt21)
            aux = cv.gt21_enum_from_label(b1 + b2)
        elif cat == 2:
            b1, b2 = cv.hetero_SNP_bases_from(gt21)
            aux = cv.gt21_enum_from_label(b1 + b2)
        elif cat == 3:
            l1 = calls["ins"]
        elif cat == 4:
            l1 = calls["ins"]
            aux = "ACGT".index(lists[7][lists[9].index(maximum)])
        elif cat == 5:
            l1, l2 = calls["ins_pysam"], calls["ins"]
        elif cat == 6:
            l1 = calls["del"]
        elif cat == 7:
            l1 = calls["del"]
            aux = "ACGT".index(lists[14][lists[16].index(maximum)])
        elif cat == 8:
            l2 = calls["del"]
            # variant_length_1 is only visible through the second allele: ref[0] + ref[l1+1:]
            second = alt_out.split(",")[1]
            l1 = len(ref_out) - len(second)
        elif cat == 9:
            l1, l2 = calls["del"], calls["ins"]
        dec[k] = (cat, l1, l2, aux)
        maxp[k] = np.float32(maximum)
        depth[k] = np.float32(sum(X[k][16, :, cv.Channel.delete] + X[k][16, :, cv.Channel.reference]))
    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, "decision_cases.npz"), probs=P, ref_bases=ref_bases, X16=X[:, 16].copy(),
                        decision=dec, max_probability=maxp, read_depth=depth)
    print("decision_cases.npz: %d cases, categories seen: %s" % (n, np.bincount(dec[:, 0], minlength=10)))


if __name__ == "__main__":
    main()
