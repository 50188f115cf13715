"""First-choice variant decision.

CPU: the oracle restatement against the golden outcomes the REFERENCE's own output_from produced
(tests/golden/decision_cases.npz, oracle/gen_golden_decision.py) - parity pinned.
GPU (-m gpu): the decide_sites kernel through the C-ABI against the same goldens (bit-exact), against the oracle on
the model's real outputs, and the fused clairb_predict_decide path against the two-call path.
"""
import os

import numpy as np
import pytest

from clair_b200 import _lib, decision, synth
from oracle import decision_oracle as D

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "decision_cases.npz")


@pytest.fixture(scope="module")
def cases():
    with np.load(GOLD) as z:
        return {k: z[k] for k in z.files}


def golden_X(cases):
    X = np.zeros((len(cases["probs"]), 33, 8, 4), np.float32)
    X[:, 16] = cases["X16"]
    return X


def test_oracle_matches_the_reference_outcomes(cases):
    dec, maxp, depth = D.decide(cases["probs"], cases["ref_bases"], golden_X(cases))
    np.testing.assert_array_equal(dec, cases["decision"])
    np.testing.assert_array_equal(maxp, cases["max_probability"])          # float32, bit-exact
    np.testing.assert_array_equal(depth, cases["read_depth"])
    assert set(cases["decision"][:, 0]) == set(range(10))                  # every category is covered


def test_golden_cases_cover_ties_and_degenerate_heads(cases):
    P = cases["probs"]
    assert (P.sum(1) == 0).any()                                           # all-zero probabilities -> "reference"
    zero = np.where(P.sum(1) == 0)[0][0]
    assert cases["decision"][zero, 0] == 0
    onehot = (P.max(1) == 1.0) & ((P == 0).sum(1) >= 80)
    assert onehot.sum() >= 21


def test_outcome_list_sizes_follow_the_reference():
    p = np.full(90, 0.1, np.float32)
    lists = D.outcome_lists(p[:21], p[21:24], p[24:57], p[57:], 2)
    # Ref, homoSNP, heteroSNP, homoIns, ACGT+Ins, InsIns, homoDel, ACGT+Del, DelDel, InsDel (call_var.py:344-424)
    assert [len(x) for x in lists] == [1, 4, 6, 16, 64, 256, 16, 64, 240, 512]


def test_host_helpers():
    infos = [["chr1", "10", "A" * 16 + b + "C" * 16] for b in "ACGTURYSWKMBDHVN"]
    codes = decision.ref_base_codes(infos)
    assert codes.tolist() == ["ACGT".index(c) for c in "ACGTTACCAGACAAAA"]          # shared/utils.py:19-29
    assert decision.flags_tuple(4) == (False,) * 4 + (True,) + (False,) * 5
    assert decision.snp_bases(6) == ("C", "T")
    rec = np.zeros((2, _lib.DECISION_WORDS), np.int32)
    rec[0, :4] = (5, 2, 7, 0)
    rec.view(np.float32)[0, 4:6] = (0.25, 31.0)
    rec[0, 6] = 411
    rec.view(np.float32)[0, 7] = 12.0
    d = decision.unpack(rec)
    assert (d.category[0], d.len1[0], d.len2[0], d.max_probability[0], d.read_depth[0]) == (5, 2, 7, 0.25, 31.0)
    assert (d.quality[0], d.supported_reads[0]) == (411, 12.0)


# ---- GPU ----------------------------------------------------------------------------------------------------------

@pytest.mark.gpu
def test_kernel_matches_the_reference_outcomes(gpu_model, cases):
    d = gpu_model.decide(cases["probs"], cases["ref_bases"], golden_X(cases))
    got = np.stack([d.category, d.len1, d.len2, d.aux], axis=1)
    np.testing.assert_array_equal(got, cases["decision"])
    np.testing.assert_array_equal(d.max_probability, cases["max_probability"])
    np.testing.assert_array_equal(d.read_depth, cases["read_depth"])
    d2 = gpu_model.decide(cases["probs"], cases["ref_bases"])             # no tensor: depth reported as 0
    np.testing.assert_array_equal(d2.category, d.category)
    assert (d2.read_depth == 0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("n,dtype", [(1, np.float32), (777, np.float32), (2500, np.int16)])
def test_fused_predict_decide_equals_two_calls_and_oracle(gpu_model, n, dtype):
    counts = synth.synthetic_counts(n, seed=400 + n).astype(np.int32)
    counts[..., 1:] -= counts[..., 0:1]
    X = counts.astype(dtype)
    ref = np.random.default_rng(n).integers(0, 4, n).astype(np.uint8)
    probs = gpu_model.predict_packed(X)
    pred, d = gpu_model.predict_and_decide(X, ref)
    np.testing.assert_array_equal(np.concatenate(pred, axis=1), probs)     # same forward, bit for bit
    assert gpu_model.prediction is pred
    two = gpu_model.decide(probs, ref, X)
    for a, b in zip(d, two):
        np.testing.assert_array_equal(a, b)
    sel = np.arange(0, n, max(1, n // 300))
    dec, maxp, depth = D.decide(probs[sel], ref[sel], X[sel].astype(np.float32))
    np.testing.assert_array_equal(np.stack([d.category, d.len1, d.len2, d.aux], axis=1)[sel], dec)
    np.testing.assert_array_equal(d.max_probability[sel], maxp)
    np.testing.assert_array_equal(d.read_depth[sel], depth)


@pytest.mark.gpu
def test_decision_multi_chunk_pipeline(weights1234, monkeypatch):
    from clair_b200.model import Clair
    monkeypatch.setenv("CLAIRB_CHUNK_SITES", "1024")
    m = Clair(max_sites=4096, batch_sites=1000)
    monkeypatch.delenv("CLAIRB_CHUNK_SITES")
    m.set_weights(weights1234)
    X = synth.synthetic_tensors(3300, seed=5)
    ref = (np.arange(3300) % 4).astype(np.uint8)
    pred, d = m.predict_and_decide(X, ref)
    probs = np.concatenate(pred, axis=1)
    np.testing.assert_array_equal(probs, m.predict_packed(X))
    dec, maxp, depth = D.decide(probs[::11], ref[::11], X[::11])
    np.testing.assert_array_equal(np.stack([d.category, d.len1, d.len2, d.aux], axis=1)[::11], dec)
    np.testing.assert_array_equal(d.max_probability[::11], maxp)
    np.testing.assert_array_equal(d.read_depth[::11], depth)
    with pytest.raises(ValueError):
        m.predict_and_decide(X, ref[:10])
    with pytest.raises(ValueError):
        m.decide(probs, np.full(3300, 7, np.uint8))
    m.close()


@pytest.mark.gpu
def test_kernel_tie_breaking_on_quantised_probabilities(gpu_model):
    # probabilities restricted to a few powers of two make thousands of outcome products exactly equal: the winner is then
    # decided purely by the reference's tie order (category chain, then list.index), which the oracle reproduces from the
    # reference-pinned goldens and the kernel must reproduce from the oracle
    rng = np.random.default_rng(99)
    n = 600
    levels = np.array([0.0, 0.125, 0.25, 0.5], np.float32)
    P = levels[rng.integers(0, 4, size=(n, 90))]
    P[:50, 24:90] = 0.25                                         # whole length heads equal
    P[50:100, 0:21] = 0.5                                        # whole gt21 head equal
    ref = rng.integers(0, 4, n).astype(np.uint8)
    d = gpu_model.decide(P, ref)
    dec, maxp, _ = D.decide(P, ref)
    np.testing.assert_array_equal(np.stack([d.category, d.len1, d.len2, d.aux], axis=1), dec)
    np.testing.assert_array_equal(d.max_probability, maxp)
    assert len(set(dec[:, 0])) >= 6


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="runs the reference's own batch_output; only where /root/reference exists")
def test_first_choice_inside_the_reference_output_stage():
    """decision.FirstChoice installed as call_var.output_from: the reference's batch_output prints the same VCF rows, with the
    reference / SNP sites answered from the decision records (here from the decision oracle; on a GPU from predict_and_decide)."""
    import time
    from clair_b200 import decision, synth
    from oracle import decision_oracle as DO
    from oracle import gen_golden_decision as GD
    cv = GD.import_reference_call_var()
    rng = np.random.default_rng(8)
    n = 400
    x = synth.synthetic_tensors(n, seed=33)
    infos = [["chr2", str(500 + 40 * i), "".join(rng.choice(list("ACGT"), size=33))] for i in range(n)]
    heads = []
    for i in range(n):                                     # a call set: mostly reference / SNP sites, some indels
        z = [rng.normal(0, 1.5, k) for k in (21, 3, 33, 33)]
        if i % 8:
            z[0][rng.integers(0, 10)] += 7                 # an ACGT-pair label
            z[2][16] += 6                                  # both lengths zero
            z[3][16] += 6
        heads.append(np.concatenate([GD.softmax(v) for v in z]))
    P = np.stack(heads)
    batch_Y = [P[:, 0:21], P[:, 21:24], P[:, 24:57], P[:, 57:90]]
    records, max_p, depth = DO.decide(P, decision.ref_base_codes(infos), x)
    dec = decision.Decision(records[:, 0], records[:, 1], records[:, 2], records[:, 3], max_p, depth)
    config = cv.OutputConfig(is_show_reference=True, is_debug=False, is_haploid_precision_mode_enabled=False,
                             is_haploid_sensitive_mode_enabled=False, is_output_for_ensemble=False, quality_score_for_pass=None)

    class Stingy(GD.Recorder):
        """indel-base helpers that come back empty for some lengths: the reference's loop then moves on to its next candidate"""

        def insertion_bases_using(self, tensor_input, variant_length, contig, position):
            return ("", 0) if variant_length % 3 == 0 else GD.Recorder.insertion_bases_using(self, tensor_input, variant_length, contig, position)

        def deletion_bases_using(self, tensor_input, variant_length, contig, position, reference_sequence):
            return ("", 0) if variant_length % 4 == 0 else GD.Recorder.deletion_bases_using(self, tensor_input, variant_length, contig, position, reference_sequence)

    def run(helpers):
        lines, rec = [], helpers()
        util = cv.OutputUtilities(print_debug_message=lambda *a: lines.append(("debug", a[0], a[1], a[-1])),
                                  insertion_bases_using=rec.insertion_bases_using, deletion_bases_using=rec.deletion_bases_using,
                                  insertion_bases_using_pysam_using=rec.insertion_bases_using_pysam_using,
                                  output=lines.append, output_header=lambda: None, close_opened_files=lambda: None)
        t = time.perf_counter()
        cv.batch_output((x, infos), batch_Y, config, util)
        return lines, time.perf_counter() - t

    original = cv.output_from
    for helpers in (GD.Recorder, Stingy):
        want, t_reference = run(helpers)
        first_choice = decision.FirstChoice(original)
        first_choice.load(infos, dec)
        cv.output_from = first_choice
        try:
            got, t_first_choice = run(helpers)
        finally:
            cv.output_from = original
        assert got == want and len(want) > n // 2
        total = first_choice.served + first_choice.deferred
        assert total > n // 2
        if helpers is GD.Recorder:
            assert first_choice.deferred <= 2                  # only InsIns / DelDel first choices whose two alleles coincide
            assert len({line.split("\t")[4].count(",") for line in want if isinstance(line, str)}) == 2     # single and multi-allelic rows
        else:
            assert 0 < first_choice.deferred < total // 4      # the empty-handed helpers send their sites to the reference's loop
    print("batch_output: %.1f ms/site with the reference's output_from, %.2f ms/site with FirstChoice (%d of %d sites from records)"
          % (1e3 * t_reference / n, 1e3 * t_first_choice / n, first_choice.served, first_choice.served + first_choice.deferred))
