"""Batched VCF output stage behind the forward: what the reference's `batch_output` ->
`output_with` prints per site (clair/call_var.py:1002-1236), a batch at a time.

The reference spends ~1 ms of Python per site here (1.2 k float products, `in` tests over lists, string work): three
orders of magnitude slower than the forward in front of it.  With the decision records of `Clair.predict_and_decide` in
hand - category, variant lengths, gt21 label, read depth, quality score and supporting-read count of every site, computed
on the device right behind the heads - the stage becomes:

  * reference and SNP calls (the bulk of every call set): numpy over the whole batch, rows formatted by the C-ABI
    library (`clairb_format_vcf_rows`), one `output_utilities.output` call per run of consecutive rows;
  * insertion / deletion calls: REF / ALT through the same indel-base helpers of `output_utilities` the reference calls
    for its first choice (`decision.FirstChoice`), numbers from the record;
  * sites whose first choice does not stand (a helper came back empty: the reference's loop moves on to its next
    candidate) go to `fallback`, the reference's own `output_with`, if the host program supplied it.

Usage, in place of `call_var.batch_output`:
    stage = output.BatchOutput(output_config, output_utilities, fallback=call_var.output_with)
    call_var.run_batches(m, tensor_generator, stage, with_decision=True)

Numeric semantics are those of the reference's pinned numpy 1.18 (README.md:127): float32 product for the probability of
the call, float64 for everything a Python float touches (`1.0 - p`, `(supported + 0.0) / read_depth`).
`is_output_for_ensemble` (batch_output_for_ensemble) is not covered.
"""
import ctypes
from math import e, log

import numpy as np

from . import _lib, decision as _decision

CENTER = 16
_GENOTYPES = ("0/0", "1/1", "0/1", "1/2")                         # clair/task/genotype.py:3
_ACGT_CODE = np.full(256, 255, np.uint8)
for _i, _b in enumerate(b"ACGT"):
    _ACGT_CODE[_b] = _i
_BASIC = np.zeros(256, bool)
for _b in b"ACGTU":                                               # shared/utils.py:31 BASIC_BASES
    _BASIC[_b] = True
_BASE_CHAR = np.frombuffer(b"ACGT", np.uint8)
# label bases of the gt21 pair labels 0..9: AA AC AG AT CC CG CT GG GT TT (clair/task/gt21.py:3-13)
_LABEL_1 = np.array([0, 0, 0, 0, 1, 1, 1, 2, 2, 3], np.uint8)
_LABEL_2 = np.array([0, 1, 2, 3, 1, 2, 3, 2, 3, 3], np.uint8)
_BASE2NUM = dict(zip("ACGT", (0, 1, 2, 3)))
_REFERENCE, _INSERT, _DELETE, _SNP = 0, 1, 2, 3                    # Channel, clair/call_var.py:32-37
_GT21_LABELS = _decision.GT21_LABELS


class NeedsFallback(RuntimeError):
    """A site whose first choice does not stand and no `fallback` (the reference's output_with) was supplied."""


def _partial_label(ref, alt):                                     # clair/task/gt21.py:64-69
    return "Del" if len(ref) > len(alt) else "Ins" if len(ref) < len(alt) else alt[0]


def _mix_labels(l1, l2):                                          # clair/task/gt21.py:72-89
    if len(l1) == 1 and len(l2) == 1:
        return l1 + l2 if l1 <= l2 else l2 + l1
    t1, t2 = (l2, l1) if (len(l1) > 1 and len(l2) == 1) else (l1, l2)
    if len(t2) > 1 and len(t1) == 1:
        return t1 + t2
    if l1 and l2 and l1 == l2:
        return l1 + l2
    return "InsDel"


def quality_score(reference, alternate, genotype_string, gt21_probabilities, genotype_probabilities):
    """clair/call_var.py:568-586 (used where the device record's score does not apply)."""
    g1, g2 = int(genotype_string[0]), int(genotype_string[2])
    arr = alternate.split(",")
    if len(arr) == 1:                                             # clair/task/gt21.py:98-103
        arr = [reference if g1 == 0 or g2 == 0 else arr[0]] + arr
    gt21 = _GT21_LABELS.index(_mix_labels(_partial_label(reference, arr[0]), _partial_label(reference, arr[1])))
    genotype = 0 if (g1 == 0 and g2 == 0) else 1 if g1 == g2 else 2    # genotype.py:20-33 (multi folded into hetero)
    p = float(np.float32(gt21_probabilities[gt21]) * np.float32(genotype_probabilities[genotype]))
    tmp = max((-10 * log(e, 10)) * log(((1.0 - p) + 1e-300) / (p + 1e-300)) + 16, 0)
    return int(round(tmp * tmp))


def supported_reads(x, flags, reference_base, alternate_base):
    """clair/call_var.py:1087-1151 for one site (used where the device record's count does not apply)."""
    x = np.asarray(x, dtype=np.float64)
    c, multi = CENTER, "," in alternate_base
    snp = lambda b: (x[c, _BASE2NUM[b], _SNP] + x[c, _BASE2NUM[b] + 4, _SNP] + x[c, _BASE2NUM[b], _REFERENCE] +
                     x[c, _BASE2NUM[b] + 4, _REFERENCE])
    if flags[0]:
        return float(x[c, _BASE2NUM[reference_base], _REFERENCE] + x[c, _BASE2NUM[reference_base] + 4, _REFERENCE])
    if flags[1] or flags[2]:
        return float(sum(snp(b) for b in alternate_base if b != ","))
    ins, dele, snp17 = x[c + 1, :, _INSERT].sum(), x[c + 1, :, _DELETE].sum(), x[c + 1, :, _SNP].sum()
    if flags[3] or flags[5]:
        return float(ins - snp17)
    if flags[4]:
        return float((ins - snp17) + (snp(alternate_base.split(",")[0][0]) if multi else 0))
    if flags[6] or flags[8]:
        return float(dele)
    if flags[7]:
        return float(dele + (snp(alternate_base.split(",")[1][0]) if multi else 0))
    return float(ins + dele - snp17)


class BatchOutput(object):
    """Output stage for `call_var.run_batches(..., with_decision=True)`: called as stage(mini_batch, prediction, decision)."""

    def __init__(self, output_config, output_utilities, fallback=None):
        if getattr(output_config, "is_output_for_ensemble", False):
            raise ValueError("is_output_for_ensemble is not covered by BatchOutput (batch_output_for_ensemble)")
        self.config, self.util, self.fallback = output_config, output_utilities, fallback
        self.fast_rows = self.slow_rows = self.fallback_sites = 0
        self._lib = _lib.load()

    # ---- reference / SNP calls, a batch at a time ------------------------------------------------------------------
    def _fast_rows(self, infos, idx, dec, centre):
        """Formatted rows of the fast sites `idx` that print anything -> (kept site indices, blob, row_end)."""
        cfg = self.config
        cat = np.asarray(dec.category)[idx]
        aux = np.asarray(dec.aux)[idx]
        ref = _ACGT_CODE[centre[idx]]
        l1, l2 = _LABEL_1[aux], _LABEL_2[aux]
        multi = (cat == 2) & (l1 != ref) & (l2 != ref)
        alt1 = np.where(cat == 0, ref, np.where(multi | (l1 != ref), l1, l2))
        keep = np.where(cat == 0, bool(cfg.is_show_reference), alt1 != ref)                        # :1046-1050
        if cfg.is_haploid_precision_mode_enabled:                                                   # :1066-1071
            keep &= cat != 2
        elif cfg.is_haploid_sensitive_mode_enabled:                                                 # :1072-1074
            keep &= ~multi
        sel = np.flatnonzero(keep)
        if sel.size == 0:
            return idx[sel], b"", np.empty(0, np.int64)
        idx, cat, ref, l2, alt1, multi = idx[sel], cat[sel], ref[sel], l2[sel], alt1[sel], multi[sel]
        n = idx.size
        gt = np.where(multi, 3, np.where(cat == 0, 0, np.where(cat == 1, 1, 2))).astype(np.uint8)   # :1076-1085
        if cfg.is_haploid_precision_mode_enabled or cfg.is_haploid_sensitive_mode_enabled:          # :1164-1166
            gt = np.where(gt == 0, 4, 5).astype(np.uint8)
        alt = np.zeros((n, 4), np.uint8)
        alt[:, 0] = _BASE_CHAR[alt1]
        alt[multi, 1] = ord(",")
        alt[multi, 2] = _BASE_CHAR[l2[multi]]
        depth = np.asarray(dec.read_depth)[idx].astype(np.float64)
        support = np.asarray(dec.supported_reads)[idx].astype(np.float64)
        af = np.minimum(support / depth, 1.0)                                                       # :1152-1154
        quality = np.ascontiguousarray(np.asarray(dec.quality)[idx], dtype=np.int32)
        if cfg.quality_score_for_pass is None:                                                      # :70-75
            filt = np.zeros(n, np.uint8)
        else:
            filt = np.where(quality >= cfg.quality_score_for_pass, 1, 2).astype(np.uint8)
        contigs = [info[0] for info in infos]
        if contigs.count(contigs[0]) == len(contigs):                 # one contig per batch is the rule
            name = contigs[0].encode()
            blob_in = name * n
            ctg_off = (np.arange(n + 1, dtype=np.int64) * len(name)).astype(np.int32)
        else:
            names = [contigs[i].encode() for i in idx.tolist()]
            ctg_off = np.zeros(n + 1, np.int32)
            np.cumsum([len(s) for s in names], out=ctg_off[1:])
            blob_in = b"".join(names)
        pos = np.fromstring(" ".join(info[1] for info in infos), dtype=np.int64, sep=" ")       # one C-level parse per batch
        if pos.shape[0] != len(infos):
            raise ValueError("positions of the batch are not plain integers")
        pos = np.ascontiguousarray(pos[idx])
        refc = np.ascontiguousarray(_BASE_CHAR[ref])
        depth_i = np.ascontiguousarray(depth.astype(np.int32))                                      # "%d" of a float truncates
        af = np.ascontiguousarray(af, dtype=np.float64)
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        need = ctypes.c_int64()
        cap = len(blob_in) + 128 * n + 1                  # every field of a row but the contig is bounded: one call, no measuring pass
        out = ctypes.create_string_buffer(cap)
        row_end = np.empty(n, np.int64)
        rc = self._lib.clairb_format_vcf_rows(n, blob_in, p(ctg_off), p(pos), p(refc), p(alt), p(quality), p(filt), p(gt), p(depth_i),
                                              p(af), out, cap, ctypes.byref(need), p(row_end))
        _lib.check(rc, None, "clairb_format_vcf_rows")
        return idx, out.raw[:need.value], row_end

    # ---- everything else, a site at a time ------------------------------------------------------------------------
    def _slow_site(self, i, X, infos, prediction, dec, centre):
        cfg, util = self.config, self.util
        chromosome, position, sequence = infos[i]
        position = int(position)
        probs = [a[i] for a in prediction]
        depth = float(dec.read_depth[i])
        if cfg.is_debug:
            return self._to_fallback(i, X, infos, probs, "debug mode prints through the reference's own output_with")
        if depth == 0:                                                                              # :1019-1029
            util.print_debug_message(chromosome, position, probs[0], probs[1], probs[2], probs[3], "Read Depth is zero")
            return
        category = int(dec.category[i])
        x = X[i]
        answer = _decision.FirstChoice._first_choice(x, sequence, chromosome, position, CENTER, util, category,
                                                     int(dec.len1[i]), int(dec.len2[i]), int(dec.aux[i]))
        if answer is None:
            return self._to_fallback(i, X, infos, probs, "its first choice does not stand")
        flags, (reference_base, alternate_base) = answer
        if (not cfg.is_show_reference and flags[0]) or (not flags[0] and reference_base == alternate_base):   # :1046-1050
            return
        multi = "," in alternate_base
        hetero = flags[2] or flags[4] or flags[5] or flags[7] or flags[8]
        if cfg.is_haploid_precision_mode_enabled and (hetero or flags[9]):                          # :1066-1071
            return
        if not cfg.is_haploid_precision_mode_enabled and cfg.is_haploid_sensitive_mode_enabled and multi:
            return
        genotype_string = _GENOTYPES[3] if multi else _GENOTYPES[0] if flags[0] else \
            _GENOTYPES[1] if (flags[1] or flags[3] or flags[6]) else _GENOTYPES[2] if hetero else ""
        if _ACGT_CODE[centre[i]] != 255:
            # the record's numbers assume exactly this first choice with an A/C/G/T reference base
            support, quality = float(dec.supported_reads[i]), int(dec.quality[i])
        else:
            support = supported_reads(x, flags, reference_base, alternate_base)
            quality = quality_score(reference_base, alternate_base, genotype_string, probs[0], probs[1])
        allele_frequency = min((support + 0.0) / depth, 1)                                          # :1152-1154
        if cfg.is_haploid_precision_mode_enabled or cfg.is_haploid_sensitive_mode_enabled:          # :1164-1166
            genotype_string = "1" if "1" in genotype_string else "0"
        filtration = "." if cfg.quality_score_for_pass is None else "PASS" if quality >= cfg.quality_score_for_pass else "LowQual"
        self.slow_rows += 1
        util.output("%s\t%d\t.\t%s\t%s\t%d\t%s\t%s\tGT:GQ:DP:AF\t%s:%d:%d:%.4f" % (
            chromosome, position, reference_base, alternate_base, quality, filtration, ".", genotype_string, quality, depth,
            allele_frequency))

    def _to_fallback(self, i, X, infos, probs, why):
        self.fallback_sites += 1
        if self.fallback is None:
            raise NeedsFallback("site %s:%s: %s; pass fallback=<the reference's output_with> to BatchOutput" % (infos[i][0], infos[i][1], why))
        self.fallback(X[i], infos[i], probs[0], probs[1], probs[2], probs[3], self.config, self.util)

    def __call__(self, mini_batch, prediction, dec):
        X, infos = mini_batch
        n = len(infos)
        if len(prediction[0]) != n:                                                                 # :1206-1210
            raise ValueError("Inconsistent shape between input tensor and output predictions %d/%d" % (n, len(prediction[0])))
        centre = np.frombuffer("".join(info[2][CENTER] for info in infos).encode("latin-1"), np.uint8)
        basic = _BASIC[centre]                                                                      # :1012-1013
        cat = np.asarray(dec.category)
        fast = basic & (_ACGT_CODE[centre] != 255) & (cat <= 2) & (np.asarray(dec.read_depth) != 0) & (not self.config.is_debug)
        slow = np.flatnonzero(basic & ~fast)
        kept, blob, row_end = self._fast_rows(infos, np.flatnonzero(fast), dec, centre)
        self.fast_rows += int(kept.size)
        if slow.size == 0:
            if kept.size:
                self.util.output(blob.decode())
            return
        # rows leave in site order: runs of fast rows between the slow sites go out as one call each
        at = np.searchsorted(kept, slow)                 # number of fast rows in front of every slow site
        done = 0
        for j, i in enumerate(slow.tolist()):
            k = int(at[j])
            if k > done:
                self.util.output(blob[(int(row_end[done - 1]) + 1 if done else 0):int(row_end[k - 1])].decode())
                done = k
            self._slow_site(i, X, infos, prediction, dec, centre)
        if done < kept.size:
            self.util.output(blob[(int(row_end[done - 1]) + 1 if done else 0):].decode())
