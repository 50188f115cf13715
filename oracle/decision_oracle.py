"""CPU restatement of the first-choice variant decision of the reference's VCF stage.  TEST INFRASTRUCTURE ONLY
(imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg; never by the product path).

Follows, line by line, the part of `clair/call_var.py` that turns the four probability vectors of one site into
"which outcome wins":
    possible_outcome_probabilites_from   clair/call_var.py:589-690   (the ~1.2 k outcome products)
    *_tuples_from                        clair/call_var.py:344-424   (list orders and the length tuples)
    output_from, first pass of the loop  clair/call_var.py:732-760   (max, then the is_* tests in elif order)
    homo/hetero_SNP_bases_from           clair/call_var.py:60-67     (arg-max over the SNP subsets of gt21)
    read depth                           clair/call_var.py:1021-1024
Arithmetic is float32 throughout, exactly as the reference computes it: the probabilities are numpy.float32 scalars
(rows of the float32 arrays `predict` returns) and numpy keeps float32 * float32 in float32; products associate left to
right as written in the reference.  Ties resolve as in the reference: the first category in elif order that contains
the maximum, and `list.index` (first occurrence) inside it.

PARITY PINNED: tests/golden/decision_cases.npz holds the outcome the reference's own `output_from` produced for every
case (oracle/gen_golden_decision.py imports /root/reference/clair/call_var.py with pysam / clair.model stubbed).

Decision record per site (int32 x 4 + two float32):
    category  0 reference, 1 homo SNP, 2 hetero SNP, 3 homo Ins, 4 hetero ACGT+Ins, 5 hetero InsIns, 6 homo Del,
              7 hetero ACGT+Del, 8 hetero DelDel, 9 InsDel            (order of the flags tuple, call_var.py:931-937)
    len1,len2 SNP: 0,0.  homo Ins/Del: length,0.  ACGT+Ins/Del: length,0.  InsIns/DelDel: the (min,max) tuple.
              InsDel: the tuple stored by hetero_InsDel_tuples_from (deletion length, insertion length).
    aux       reference: gt21 index of the reference base pair.  SNP: gt21 index of the winning label (arg-max over
              the subset).  ACGT+Ins/Del: hetero base 0..3 (ACGT).  otherwise 0.
    max_probability, read_depth
"""
import numpy as np

INDEX_OFFSET = 16                     # clair/task/variant_length.py:6
VL_MAX = 16
HOMO_SNP_GT21 = (0, 4, 7, 9)          # AA CC GG TT        clair/task/gt21.py:111
HETERO_SNP_GT21 = (1, 2, 3, 5, 6, 8)  # AC AG AT CG CT GT  clair/task/gt21.py:114
GT_DELDEL, GT_ADEL, GT_INSINS, GT_AINS, GT_INSDEL = 10, 11, 15, 16, 20
REF_GT21 = (0, 4, 7, 9)               # gt21_enum_from_label(base + base) for A, C, G, T
HOMO_REF, HOMO_VAR, HETERO_VAR = 0, 1, 2   # clair/task/genotype.py:6-10

CATEGORIES = ("reference", "homo_SNP", "hetero_SNP", "homo_insertion", "hetero_ACGT_Ins", "hetero_InsIns",
              "homo_deletion", "hetero_ACGT_Del", "hetero_DelDel", "insertion_and_deletion")

F = np.float32


def outcome_lists(gt21, geno, vl1, vl2, ref_base):
    """The ten outcome lists of possible_outcome_probabilites_from as [(probability, len1, len2, aux)], in the
    reference's list order.  All inputs float32 vectors; ref_base 0..3."""
    gt21, geno, vl1, vl2 = (np.asarray(a, dtype=F) for a in (gt21, geno, vl1, vl2))
    homo_ref, homo_var, het_var = geno[HOMO_REF], geno[HOMO_VAR], geno[HETERO_VAR]
    O = INDEX_OFFSET
    vl0 = vl1[O] * vl2[O]                                                      # :600-603
    lists = []
    lists.append([(vl0 * homo_ref * gt21[REF_GT21[ref_base]], 0, 0, REF_GT21[ref_base])])            # :606-608
    # the SNP lists keep the subset order; the label the reference prints is the arg-max over the subset (:60-67),
    # i.e. the first maximum - carried as aux of every entry so that whichever entry wins reports the same label
    homo_best = HOMO_SNP_GT21[int(np.argmax([gt21[g] for g in HOMO_SNP_GT21]))]
    het_best = HETERO_SNP_GT21[int(np.argmax([gt21[g] for g in HETERO_SNP_GT21]))]
    lists.append([(vl0 * homo_var * gt21[g], 0, 0, homo_best) for g in HOMO_SNP_GT21])               # :610-612
    lists.append([(vl0 * het_var * gt21[g], 0, 0, het_best) for g in HETERO_SNP_GT21])               # :613-615
    extra = homo_var * gt21[GT_INSINS]                                                                 # :620
    homo_ins = [(vl1[i + O] * vl2[i + O] * extra, i, 0, 0) for i in range(1, VL_MAX + 1)]             # :344-349
    extra = het_var * gt21[GT_INSINS]                                                                  # :625
    insins = [(vl1[i + O] * vl2[j + O] * extra, min(i, j), max(i, j), 0)
              for i in range(1, VL_MAX + 1) for j in range(1, VL_MAX + 1)]                            # :364-374
    acgt_ins = []
    for i in range(1, VL_MAX + 1):                                                                     # :352-361, :632-639
        p = max(vl1[O] * vl2[i + O], vl1[i + O] * vl2[O])
        for b in range(4):
            acgt_ins.append((p * gt21[GT_AINS + b] * het_var, i, 0, b))
    extra = homo_var * gt21[GT_DELDEL]                                                                 # :648
    homo_del = [(vl1[-i + O] * vl2[-i + O] * extra, i, 0, 0) for i in range(1, VL_MAX + 1)]           # :377-382
    extra = het_var * gt21[GT_DELDEL]                                                                  # :653
    deldel = [(vl1[-i + O] * vl2[-j + O] * extra, min(i, j), max(i, j), 0)
              for i in range(1, VL_MAX + 1) for j in range(1, VL_MAX + 1) if i != j]                  # :397-408
    acgt_del = []
    for i in range(1, VL_MAX + 1):                                                                     # :385-394, :660-667
        p = max(vl1[O] * vl2[-i + O], vl1[-i + O] * vl2[O])
        for b in range(4):
            acgt_del.append((p * gt21[GT_ADEL + b] * het_var, i, 0, b))
    extra = het_var * gt21[GT_INSDEL]                                                                  # :676
    insdel = []
    for i in range(1, VL_MAX + 1):                                                                     # :411-424
        for j in range(1, VL_MAX + 1):
            insdel.append((vl1[i + O] * vl2[-j + O] * extra, j, i, 0))
            insdel.append((vl1[-i + O] * vl2[j + O] * extra, i, j, 0))
    # order of the flags tuple / elif chain (:750-758): Ref, homoSNP, heteroSNP, homoIns, ACGT+Ins, InsIns, homoDel,
    # ACGT+Del, DelDel, InsDel
    lists += [homo_ins, acgt_ins, insins, homo_del, acgt_del, deldel, insdel]
    return lists


def decide_site(probs90, ref_base):
    """(category, len1, len2, aux, max_probability) of one site, first pass of output_from's loop (:732-760)."""
    p = np.asarray(probs90, dtype=F)
    lists = outcome_lists(p[0:21], p[21:24], p[24:57], p[57:90], int(ref_base))
    maximum = max(max(e[0] for e in lst) for lst in lists)                     # :734-745
    for cat, lst in enumerate(lists):                                          # elif order, list.index = first hit
        for e in lst:
            if e[0] == maximum:
                return cat, e[1], e[2], e[3], F(maximum)
    raise AssertionError("unreachable")


def read_depth(x):
    """sum(x[16,:,delete] + x[16,:,reference])  (call_var.py:1021-1024); x is one [33,8,4] tensor."""
    x = np.asarray(x, dtype=F)
    return F(sum(x[16, :, 2] + x[16, :, 0]))


def decide(probs, ref_bases, X=None):
    """Batch form: probs [n,90], ref_bases [n] in 0..3 -> (dec int32 [n,4], maxp float32 [n], depth float32 [n])."""
    n = len(probs)
    dec = np.zeros((n, 4), np.int32)
    maxp = np.zeros(n, F)
    depth = np.zeros(n, F)
    for i in range(n):
        c, l1, l2, aux, m = decide_site(probs[i], ref_bases[i])
        dec[i] = (c, l1, l2, aux)
        maxp[i] = m
        if X is not None:
            depth[i] = read_depth(X[i])
    return dec, maxp, depth


# ---- what output_with derives from the first choice without any string in hand (record words 6, 7) --------------------
# quality_score_from (clair/call_var.py:568-586) needs the gt21 label and genotype of the call; gt21_enum_from
# (clair/task/gt21.py:92-108) reads them off the REF / ALT / genotype strings, but for a first choice that stands they follow
# from the category: reference -> (ref ref, 0/0); SNPs -> (the label, 1/1 | 0/1 | 1/2); homo Ins / InsIns -> InsIns;
# ACGT+Ins -> <base>Ins; homo Del / DelDel -> DelDel; ACGT+Del -> <base>Del; Ins+Del -> InsDel; two-allele calls score with
# the hetero genotype (genotype_enum_for_task, clair/task/genotype.py:30-33).  tests/test_output.py pins this against the
# rows the reference's own output_with printed.
QUALITY_GT21 = {3: GT_INSINS, 5: GT_INSINS, 6: GT_DELDEL, 8: GT_DELDEL, 9: GT_INSDEL}
QUALITY_GENOTYPE = (HOMO_REF, HOMO_VAR, HETERO_VAR, HOMO_VAR, HETERO_VAR, HETERO_VAR, HOMO_VAR, HETERO_VAR, HETERO_VAR, HETERO_VAR)
LABEL_BASES = ((0, 0), (0, 1), (0, 2), (0, 3), (1, 1), (1, 2), (1, 3), (2, 2), (2, 3), (3, 3))    # AA AC AG AT CC CG CT GG GT TT


def quality_of_first_choice(probs90, category, aux):
    """int(round(max(-10 log10(e) ln((1-p+1e-300)/(p+1e-300)) + 16, 0)^2)), p the float32 product (call_var.py:568-586)."""
    from math import e, log
    p90 = np.asarray(probs90, dtype=F)
    gt21 = aux if category <= 2 else GT_AINS + aux if category == 4 else GT_ADEL + aux if category == 7 else QUALITY_GT21[category]
    p = float(p90[gt21] * p90[21 + QUALITY_GENOTYPE[category]])
    tmp = max((-10 * log(e, 10)) * log(((1.0 - p) + 1e-300) / (p + 1e-300)) + 16, 0)
    return int(round(tmp * tmp))


def support_of_first_choice(x, ref_base, category, aux):
    """Supporting-read count of the first choice (call_var.py:1087-1151); x one [33,8,4] tensor, ref_base 0..3."""
    x = np.asarray(x, dtype=np.float64)
    snp = lambda b: x[16, b, 3] + x[16, b + 4, 3] + x[16, b, 0] + x[16, b + 4, 0]
    ins, dele, snp17 = x[17, :, 1].sum(), x[17, :, 2].sum(), x[17, :, 3].sum()
    if category == 0:
        return F(x[16, ref_base, 0] + x[16, ref_base + 4, 0])
    if category == 1:
        return F(snp(LABEL_BASES[aux][0]))
    if category == 2:
        b1, b2 = LABEL_BASES[aux]
        return F(snp(b1) + snp(b2)) if (b1 != ref_base and b2 != ref_base) else F(snp(b1 if b1 != ref_base else b2))
    if category in (3, 5):
        return F(ins - snp17)
    if category == 4:
        return F((ins - snp17) + (snp(aux) if aux != ref_base else 0))
    if category in (6, 8):
        return F(dele)
    if category == 7:
        return F(dele + (snp(aux) if aux != ref_base else 0))
    return F(ins + dele - snp17)


def decide_full(probs, ref_bases, X):
    """decide() plus the two derived words: -> (dec [n,4], maxp [n], depth [n], quality int32 [n], support float32 [n])."""
    dec, maxp, depth = decide(probs, ref_bases, X)
    n = len(probs)
    quality = np.zeros(n, np.int32)
    support = np.zeros(n, F)
    for i in range(n):
        quality[i] = quality_of_first_choice(probs[i], int(dec[i, 0]), int(dec[i, 3]))
        support[i] = support_of_first_choice(X[i], int(ref_bases[i]), int(dec[i, 0]), int(dec[i, 3]))
    return dec, maxp, depth, quality, support


def categories_holding_the_maximum(probs90, ref_base):
    """How many of the ten outcome lists contain the maximum.  More than one = an exact tie between categories: the
    reference's flags tuple then carries several True entries (clair/call_var.py:750-758 tests every list), and the
    elif chains of output_with (:1076-1151) may follow a different flag than the one output_from built REF / ALT from.
    Never seen with real softmax outputs; the golden cases with probabilities quantised to powers of two are full of them."""
    p = np.asarray(probs90, dtype=F)
    lists = outcome_lists(p[0:21], p[21:24], p[24:57], p[57:90], int(ref_base))
    maximum = max(max(e[0] for e in lst) for lst in lists)
    return sum(1 for lst in lists if any(e[0] == maximum for e in lst))
