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
