"""Host side of the first-choice variant decision.

The device kernel (`csrc/decide_kernels.cuh`, entry points `clairb_predict_decide` / `clairb_decide`) answers, per
site, the question the reference answers with ~1.2 k Python float products and a chain of `in` tests: which outcome
of `possible_outcome_probabilites_from` (clair/call_var.py:589-690) is the maximum on the first pass of
`output_from`'s loop (clair/call_var.py:732-760).  This module holds the record layout and the small pieces of the
reference's bookkeeping a caller needs around it; REF/ALT string assembly, indel-base lookups and the rare retry
iterations (clair/call_var.py:762-929) stay with the caller.
"""
import collections

import numpy as np

from . import _lib

# order of the flags tuple output_from returns (clair/call_var.py:931-937)
CATEGORIES = ("reference", "homo_SNP", "hetero_SNP", "homo_insertion", "hetero_ACGT_Ins", "hetero_InsIns",
              "homo_deletion", "hetero_ACGT_Del", "hetero_DelDel", "insertion_and_deletion")
GT21_LABELS = ("AA", "AC", "AG", "AT", "CC", "CG", "CT", "GG", "GT", "TT", "DelDel", "ADel", "CDel", "GDel", "TDel",
               "InsIns", "AIns", "CIns", "GIns", "TIns", "InsDel")                       # clair/task/gt21.py:3-25
# shared/utils.py:19-29
IUPAC_TO_ACGT = dict(zip("ACGTURYSWKMBDHVN", "ACGTTACCAGACAAAA"))
BASIC_BASES = frozenset("ACGTU")                                                         # shared/utils.py:31

Decision = collections.namedtuple("Decision", "category len1 len2 aux max_probability read_depth")


def ref_base_codes(non_tensor_infos, centre=16):
    """[n] uint8 codes 0..3 (A C G T) of BASE2ACGT[reference_sequence[16]] (clair/call_var.py:718) from the
    (chromosome, position, sequence) triples the tensor generator yields."""
    codes = np.empty(len(non_tensor_infos), dtype=np.uint8)
    for i, (_, _, seq) in enumerate(non_tensor_infos):
        codes[i] = "ACGT".index(IUPAC_TO_ACGT[seq[centre]])
    return codes


def unpack(records):
    """[n,6] int32 device records -> Decision of column views into them (the two float fields are bit-cast back;
    no copies: splitting 142,000 records into six contiguous arrays costs as much as 10 % of their forward pass)."""
    r = np.ascontiguousarray(records, dtype=np.int32).reshape(-1, _lib.DECISION_WORDS)
    f = r.view(np.float32)
    return Decision(r[:, 0], r[:, 1], r[:, 2], r[:, 3], f[:, 4], f[:, 5])


def flags_tuple(category):
    """The ten booleans output_from returns for one site (clair/call_var.py:931-937)."""
    return tuple(i == int(category) for i in range(len(CATEGORIES)))


def snp_bases(aux):
    """(base1, base2) of a reference / SNP decision: the gt21 label (clair/call_var.py:60-67)."""
    label = GT21_LABELS[int(aux)]
    return label[0], label[1]


class FirstChoice(object):
    """Drop-in for the reference's `output_from` (clair/call_var.py:692-937) that answers from the device's decision records.

    `output_with` (clair/call_var.py:1002-1197) calls `output_from(x, reference_sequence, contig, position, ...)` per site and
    gets back the ten flags and (reference_base, alternate_base).  For the sites whose first choice is reference / homo SNP /
    hetero SNP - the bulk of any call set - both follow from the record alone (category + gt21 label), without building the
    ~1.2 k outcome products; every other site (indel categories, whose bases need the tensor / the BAM and may need the retry
    loop) goes to `fallback`, the reference's own function.  A maintainer installs it as

        first_choice = decision.FirstChoice(call_var.output_from); call_var.output_from = first_choice
        ... per batch:  first_choice.load(batch_chr_pos_seq, dec)      # dec from m.predict_and_decide

    Difference from the reference: when two categories tie EXACTLY for the maximum (never seen with real softmax outputs) the
    reference's tuple carries a True flag for each of them; the record knows the first in its elif order only.
    """

    def __init__(self, fallback):
        self.fallback = fallback
        self.records = {}
        self.served = self.deferred = 0

    def load(self, batch_chr_pos_seq, decision):
        category, aux = np.asarray(decision.category).tolist(), np.asarray(decision.aux).tolist()
        self.records = {(info[0], int(info[1])): (category[i], aux[i]) for i, info in enumerate(batch_chr_pos_seq)}

    def __call__(self, x, reference_sequence, contig, position, tensor_position_center, *rest, **kw):
        record = self.records.get((contig, position))
        if record is None or record[0] > 2:
            self.deferred += 1
            return self.fallback(x, reference_sequence, contig, position, tensor_position_center, *rest, **kw)
        self.served += 1
        category, aux = record
        if category == 0:                                                     # clair/call_var.py:748-753
            base = IUPAC_TO_ACGT[reference_sequence[tensor_position_center]]
            return flags_tuple(0), (base, base)
        base1, base2 = snp_bases(aux)
        reference_base = reference_sequence[tensor_position_center]
        if category == 2 and base1 != reference_base and base2 != reference_base:      # :771-776
            return flags_tuple(2), (reference_base, "{},{}".format(base1, base2))
        return flags_tuple(category), (reference_base, base1 if base1 != reference_base else base2)   # :766-769, 777-778
