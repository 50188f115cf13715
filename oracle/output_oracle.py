"""CPU restatement of the reference's per-site VCF row (`output_with`, clair/call_var.py:1002-1197) behind its
`output_from`.  TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's cpu_baseline leg; never the product path).

Follows, line by line:
    read depth, zero-depth and reference / identical-allele exits       clair/call_var.py:1016-1062
    haploid modes                                                        :1064-1074
    genotype string                                                      :1076-1085   (clair/task/genotype.py:3-17)
    supporting reads / allele frequency                                  :1087-1154
    quality score                                                        :568-586     (clair/task/gt21.py:64-108,
                                                                                       genotype.py:20-33)
    filter, row text                                                     :70-75, :1168-1197

Numeric semantics are those of the reference's pinned environment, numpy 1.18 (README.md:127): a numpy.float32 scalar
combined with a Python float or int promotes to float64 there (`1.0 - p`, `(supported + 0.0) / read_depth`), while
float32 * float32 stays float32 (the product p).  Under numpy >= 2 the same source lines would stay in float32 and
`log(0)` raises for p == 1; the restatement states the pinned behaviour explicitly in Python floats.

PARITY PINNED: tests/golden/output_rows.json.gz holds the rows the reference's own `output_with` printed
(oracle/gen_golden_output.py runs it from /root/reference with scalar semantics of numpy 1.18 restored around it).
"""
from math import e, log

import numpy as np

F = np.float32
BASE2NUM = dict(zip("ACGT", (0, 1, 2, 3)))                         # shared/utils.py:7
BASIC_BASES = frozenset("ACGTU")                                   # shared/utils.py:31
GENOTYPES = ["0/0", "1/1", "0/1", "1/2"]                           # clair/task/genotype.py:3
GT21_LABELS = ("AA", "AC", "AG", "AT", "CC", "CG", "CT", "GG", "GT", "TT", "DelDel", "ADel", "CDel", "GDel", "TDel",
               "InsIns", "AIns", "CIns", "GIns", "TIns", "InsDel")   # clair/task/gt21.py:3-25
REFERENCE, INSERT, DELETE, SNP = 0, 1, 2, 3                        # Channel, clair/call_var.py:32-37
CENTER = 16


def partial_label_from(ref, alt):                                  # clair/task/gt21.py:64-69
    if len(ref) > len(alt):
        return "Del"
    if len(ref) < len(alt):
        return "Ins"
    return alt[0]


def mix_two_partial_labels(label1, label2):                        # clair/task/gt21.py:72-89
    if len(label1) == 1 and len(label2) == 1:
        return label1 + label2 if label1 <= label2 else label2 + label1
    t1, t2 = label1, label2
    if len(label1) > 1 and len(label2) == 1:
        t1, t2 = label2, label1
    if len(t2) > 1 and len(t1) == 1:
        return t1 + t2
    if len(label1) > 0 and len(label2) > 0 and label1 == label2:
        return label1 + label2
    return "InsDel"


def gt21_enum_from(reference, alternate, genotype_1, genotype_2):  # clair/task/gt21.py:92-108
    arr = alternate.split(",")
    if len(arr) == 1:
        arr = [reference if genotype_1 == 0 or genotype_2 == 0 else arr[0]] + arr
    labels = [partial_label_from(reference, a) for a in arr]
    return GT21_LABELS.index(mix_two_partial_labels(labels[0], labels[1]))


def genotype_for_task(genotype_1, genotype_2):                     # clair/task/genotype.py:20-33
    if genotype_1 == 0 and genotype_2 == 0:
        return 0
    if genotype_1 == genotype_2:
        return 1
    return 2                                                       # hetero, and hetero_multi folded into it


def quality_score_from(reference, alternate, genotype_string, gt21_probabilities, genotype_probabilities):
    """clair/call_var.py:568-586."""
    g1, g2 = int(genotype_string[0]), int(genotype_string[2])
    gt21 = gt21_enum_from(reference, alternate, g1, g2)
    genotype = genotype_for_task(g1, g2)
    p = float(F(gt21_probabilities[gt21]) * F(genotype_probabilities[genotype]))     # float32 product, then float64
    tmp = max((-10 * log(e, 10)) * log(((1.0 - p) + 1e-300) / (p + 1e-300)) + 16, 0)
    return int(round(tmp * tmp))


def supported_reads(x, flags, reference_base, alternate_base):
    """clair/call_var.py:1087-1151; x is the [33,8,4] tensor the generator yields (channels 1..3 minus channel 0)."""
    (is_reference, is_homo_SNP, is_hetero_SNP, is_homo_insertion, is_hetero_ACGT_Ins, is_hetero_InsIns, is_homo_deletion,
     is_hetero_ACGT_Del, is_hetero_DelDel, is_insertion_and_deletion) = flags
    x = np.asarray(x, dtype=np.float64)                            # integer counts: exact in any float type
    c = CENTER
    is_multi = "," in str(alternate_base)
    snp = lambda b: (x[c, BASE2NUM[b], SNP] + x[c, BASE2NUM[b] + 4, SNP] + x[c, BASE2NUM[b], REFERENCE] + x[c, BASE2NUM[b] + 4, REFERENCE])
    count = 0
    if is_reference:
        count = x[c, BASE2NUM[reference_base], REFERENCE] + x[c, BASE2NUM[reference_base] + 4, REFERENCE]
    elif is_homo_SNP or is_hetero_SNP:
        for base in str(alternate_base):
            if base == ",":
                continue
            count += snp(base)
    elif is_homo_insertion or is_hetero_InsIns:
        count = sum(x[c + 1, :, INSERT]) - sum(x[c + 1, :, SNP])
    elif is_hetero_ACGT_Ins:
        count = (sum(x[c + 1, :, INSERT]) - sum(x[c + 1, :, SNP])) + (snp(alternate_base.split(",")[0][0]) if is_multi else 0)
    elif is_homo_deletion or is_hetero_DelDel:
        count = sum(x[c + 1, :, DELETE])
    elif is_hetero_ACGT_Del:
        count = sum(x[c + 1, :, DELETE]) + (snp(alternate_base.split(",")[1][0]) if is_multi else 0)
    elif is_insertion_and_deletion:
        count = sum(x[c + 1, :, INSERT]) + sum(x[c + 1, :, DELETE]) - sum(x[c + 1, :, SNP])
    return float(count)


def output_row(x, chr_pos_seq, gt21_probabilities, genotype_probabilities, flags, reference_base, alternate_base,
               is_show_reference=True, haploid_precision=False, haploid_sensitive=False, quality_score_for_pass=None):
    """The row output_with prints for one site given what output_from returned, or None where it prints nothing
    (non-debug mode; the zero-depth and no-base debug messages are the caller's)."""
    chromosome, position, reference_sequence = chr_pos_seq
    position = int(position)
    if reference_sequence[CENTER] not in BASIC_BASES:                              # :1012-1013
        return None
    xx = np.asarray(x, dtype=np.float64)
    read_depth = float(sum(xx[CENTER, :, DELETE] + xx[CENTER, :, REFERENCE]))       # :1016-1018
    if read_depth == 0:
        return None
    is_reference = flags[0]
    if (not is_show_reference and is_reference) or (not is_reference and reference_base == alternate_base):   # :1046-1050
        return None
    if reference_base is None or alternate_base is None:
        return None
    is_multi = "," in str(alternate_base)
    hetero = flags[2] or flags[4] or flags[5] or flags[7] or flags[8]
    if haploid_precision and (hetero or flags[9]):                                 # :1066-1071
        return None
    if haploid_sensitive and not haploid_precision and is_multi:                   # :1072-1074
        return None
    genotype_string = ""
    if is_reference:                                                               # :1076-1085
        genotype_string = GENOTYPES[0]
    elif flags[1] or flags[3] or flags[6]:
        genotype_string = GENOTYPES[1]
    elif hetero:
        genotype_string = GENOTYPES[2]
    if is_multi:
        genotype_string = GENOTYPES[3]
    count = supported_reads(x, flags, reference_base, alternate_base)
    allele_frequency = (count + 0.0) / read_depth if read_depth != 0 else 0.0      # :1152-1154
    if allele_frequency > 1:
        allele_frequency = 1
    quality_score = quality_score_from(reference_base, alternate_base, genotype_string, gt21_probabilities,
                                       genotype_probabilities)
    if haploid_precision or haploid_sensitive:                                     # :1164-1166
        genotype_string = "1" if "1" in genotype_string else "0"
    filtration = "." if quality_score_for_pass is None else ("PASS" if quality_score >= quality_score_for_pass else "LowQual")
    return "%s\t%d\t.\t%s\t%s\t%d\t%s\t%s\tGT:GQ:DP:AF\t%s:%d:%d:%.4f" % (
        chromosome, position, reference_base, alternate_base, quality_score, filtration, ".", genotype_string, quality_score,
        read_depth, allele_frequency)
