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

Decision = collections.namedtuple("Decision", "category len1 len2 aux max_probability read_depth quality supported_reads",
                                  defaults=(None, None))


def ref_base_codes(non_tensor_infos, centre=16):
    """[n] uint8 codes 0..3 (A C G T) of BASE2ACGT[reference_sequence[16]] (clair/call_var.py:718) from the
    (chromosome, position, sequence) triples the tensor generator yields."""
    codes = np.empty(len(non_tensor_infos), dtype=np.uint8)
    for i, (_, _, seq) in enumerate(non_tensor_infos):
        codes[i] = "ACGT".index(IUPAC_TO_ACGT[seq[centre]])
    return codes


def unpack(records):
    """[n,8] int32 device records -> Decision of column views into them (the float fields are bit-cast back;
    no copies: splitting 142,000 records into contiguous arrays costs as much as 10 % of their forward pass)."""
    r = np.ascontiguousarray(records, dtype=np.int32).reshape(-1, _lib.DECISION_WORDS)
    f = r.view(np.float32)
    return Decision(r[:, 0], r[:, 1], r[:, 2], r[:, 3], f[:, 4], f[:, 5], r[:, 6], f[:, 7])


def flags_tuple(category):
    """The ten booleans output_from returns for one site (clair/call_var.py:931-937)."""
    return tuple(i == int(category) for i in range(len(CATEGORIES)))


def snp_bases(aux):
    """(base1, base2) of a reference / SNP decision: the gt21 label (clair/call_var.py:60-67)."""
    label = GT21_LABELS[int(aux)]
    return label[0], label[1]


MIN_INFERRED_LENGTH, MAX_INFERRED_LENGTH = 16, 50      # clair/call_var.py:29-30 (maximum_variant_length_from, :480-484)


class FirstChoice(object):
    """Drop-in for the reference's `output_from` (clair/call_var.py:692-937) that answers from the device's decision records.

    `output_with` (clair/call_var.py:1002-1197) calls `output_from(x, reference_sequence, contig, position, ...)` per site and
    gets back the ten flags and (reference_base, alternate_base).  Both follow from the record - category, variant lengths,
    gt21 label / hetero base - without building the ~1.2 k outcome products: reference and SNP sites from the record alone,
    indel sites with the same calls to the indel-base helpers of `output_utilities` the reference makes for its first choice
    (:780-937).  Only when a helper comes back empty, where the reference's loop moves on to its next candidate, the site
    goes to `fallback`, the reference's own function.  A maintainer installs it as

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
        fields = [np.asarray(f).tolist() for f in (decision.category, decision.len1, decision.len2, decision.aux)]
        self.records = {(info[0], int(info[1])): (fields[0][i], fields[1][i], fields[2][i], fields[3][i])
                        for i, info in enumerate(batch_chr_pos_seq)}

    def __call__(self, x, reference_sequence, contig, position, tensor_position_center, gt21_probabilities,
                 genotype_probabilities, variant_length_probabilities_1, variant_length_probabilities_2, output_config,
                 output_utilities):
        record = self.records.get((contig, position))
        answer = None if record is None else self._first_choice(x, reference_sequence, contig, position,
                                                                tensor_position_center, output_utilities, *record)
        if answer is None:
            # no record, or the first choice did not yield bases: the reference's loop moves on to its next candidate
            # (clair/call_var.py:803, 819, 838-839, 853, ... `continue`), which needs its full outcome lists
            self.deferred += 1
            return self.fallback(x, reference_sequence, contig, position, tensor_position_center, gt21_probabilities,
                                 genotype_probabilities, variant_length_probabilities_1, variant_length_probabilities_2,
                                 output_config, output_utilities)
        self.served += 1
        return answer

    @staticmethod
    def _first_choice(x, reference_sequence, contig, position, center, util, category, len1, len2, aux):
        """(flags, (REF, ALT)) of the first pass of output_from's loop, or None where the reference would `continue`."""
        ref0 = reference_sequence[center]
        if category == 0:                                                     # clair/call_var.py:748-753
            base = IUPAC_TO_ACGT[ref0]
            return flags_tuple(0), (base, base)
        if category in (1, 2):                                                # :766-778
            base1, base2 = snp_bases(aux)
            if category == 2 and base1 != ref0 and base2 != ref0:
                return flags_tuple(2), (ref0, "{},{}".format(base1, base2))
            return flags_tuple(category), (ref0, base1 if base1 != ref0 else base2)
        if category in (3, 4, 5, 9):                                          # an insertion is part of the call
            insertion_bases, insertion_length = util.insertion_bases_using(
                tensor_input=x, variant_length=len2 if category in (5, 9) else len1, contig=contig, position=position)
            if category == 9:                                                 # :912-937
                deletion_bases, deletion_length = util.deletion_bases_using(
                    tensor_input=x, variant_length=len1, contig=contig, position=position, reference_sequence=reference_sequence)
                if insertion_length == 0 or deletion_length == 0:
                    return None
                ref = ref0 + deletion_bases
                return flags_tuple(9), (ref, "{},{}".format(ref[0], ref[0] + insertion_bases + ref[1:]))
            if insertion_length == 0:
                return None
            alt = ref0 + insertion_bases
            if category == 3:                                                 # :780-790
                return flags_tuple(3), (ref0, alt)
            if category == 4:                                                 # :792-810
                hetero_base = "ACGT"[aux]
                return flags_tuple(4), (ref0, "{},{}".format(hetero_base, alt) if hetero_base != ref0 else alt)
            another = util.insertion_bases_using_pysam_using(                 # :812-840
                contig=contig, position=position, minimum_insertion_length=len1,
                maximum_insertion_length=MAX_INFERRED_LENGTH if len1 >= MIN_INFERRED_LENGTH else len1,
                insertion_bases_to_ignore=insertion_bases) or insertion_bases[0:len1]
            alt_1 = ref0 + another
            return (flags_tuple(5), (ref0, "{},{}".format(alt_1, alt))) if alt_1 != alt else None
        deletion_bases, deletion_length = util.deletion_bases_using(          # categories 6, 7, 8
            tensor_input=x, variant_length=len2 if category == 8 else len1, contig=contig, position=position,
            reference_sequence=reference_sequence)
        if deletion_length == 0:
            return None
        ref = ref0 + deletion_bases
        if category == 6:                                                     # :842-857
            return flags_tuple(6), (ref, ref[0])
        if category == 7:                                                     # :859-883
            hetero_base = "ACGT"[aux]
            alt = "{},{}".format(ref[0], hetero_base + ref[1:]) if hetero_base != ref[0] else ref[0]
            return flags_tuple(7), (ref, alt)
        alt_1, alt_2 = ref[0], ref[0] + ref[len1 + 1:]                        # :885-910
        if alt_1 != alt_2 and ref != alt_1 and ref != alt_2:
            return flags_tuple(8), (ref, "{},{}".format(alt_1, alt_2))
        return None
