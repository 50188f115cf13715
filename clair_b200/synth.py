"""Synthetic pile-up tensors shaped like CreateTensor.generate_tensor output
(reference dataPrepScripts/CreateTensor.py:29-65), after the generator's
``channels 1..3 -= channel 0`` transform (clair/utils.py:96-98).  Recipe: SURVEY.md section 8d.

Channel meaning per (position, row=ACGTacgt): 0 reference-base count, 1 query-base count
(+ insertions), 2 reference-base count (+ deletions), 3 query-base count (SNP evidence).
"""
import numpy as np

T, R, C = 33, 8, 4


def synthetic_counts(n, seed=20240607, depth_lo=8, depth_hi=80):
    """Raw integer counts [n,33,8,4] (int16), before channel subtraction."""
    rng = np.random.default_rng(seed)
    depth = rng.integers(depth_lo, depth_hi + 1, size=(n, 1))
    ref_base = rng.integers(0, 4, size=(n, T))
    alt_base = (ref_base + rng.integers(1, 4, size=(n, T))) % 4
    fwd = rng.binomial(depth, 0.5, size=(n, T))            # forward-strand share of the depth
    rev = depth - fwd
    # mismatch rate: sequencing noise everywhere, a het/hom variant at the centre of ~half the sites
    rate = rng.beta(1.0, 12.0, size=(n, T)) * 0.6
    centre = rng.choice([0.0, 0.5, 1.0], size=n, p=[0.5, 0.3, 0.2])
    rate[:, T // 2] = np.maximum(rate[:, T // 2], centre * rng.uniform(0.8, 1.0, size=n))
    mis_f = rng.binomial(fwd, rate)
    mis_r = rng.binomial(rev, rate)
    dele_f = rng.binomial(fwd, 0.03)
    dele_r = rng.binomial(rev, 0.03)
    x = np.zeros((n, T, R, C), dtype=np.int32)
    ii, tt = np.meshgrid(np.arange(n), np.arange(T), indexing="ij")
    for strand, cov, mis, dele in ((0, fwd, mis_f, dele_f), (4, rev, mis_r, dele_r)):
        x[ii, tt, ref_base + strand, 0] += cov
        x[ii, tt, ref_base + strand, 2] += cov + dele
        x[ii, tt, ref_base + strand, 1] += cov - mis
        x[ii, tt, ref_base + strand, 3] += cov - mis
        x[ii, tt, alt_base + strand, 1] += mis
        x[ii, tt, alt_base + strand, 3] += mis
    ins = rng.poisson(0.4, size=(n, T, R)) * (rng.random((n, T, R)) < 0.15)
    x[..., 1] += ins
    return x.astype(np.int16)


def synthetic_tensors(n, seed=20240607, **kw):
    """float32 [n,33,8,4] as ``tensor_generator_from`` would yield them (small signed integers)."""
    x = synthetic_counts(n, seed, **kw).astype(np.float32)
    x[..., 1:] -= x[..., 0:1]
    return x


def synthetic_alignments(contig_len=1000000, depth=40, read_len=8000, ops_per_read=1001, site_spacing=50, seed=20240612):
    """A coordinate-sorted region of long noisy reads, already encoded (clair_b200.create_tensor.Alignments), its
    reference text and candidate positions - the input of CreateTensor (reference dataPrepScripts/CreateTensor.py:179-394)
    at ONT-like shape: every read alternates aligned runs with 1-3 base insertions / deletions (an op every ~8 bases).
    Built with numpy directly (a SAM text of this size would take minutes to print and parse in Python)."""
    from .create_tensor import Alignments, OP_D, OP_I, OP_M
    rng = np.random.default_rng(seed)
    n_reads = max(1, contig_len * depth // read_len)
    K = ops_per_read | 1                                        # M (I|D M)*
    n_gap = K // 2
    m_mean = max(2, (read_len - 2 * n_gap) // (n_gap + 1))
    code = np.empty((n_reads, K), np.int64)
    code[:, 0::2] = OP_M
    code[:, 1::2] = np.where(rng.random((n_reads, n_gap)) < 0.5, OP_I, OP_D)
    length = np.empty((n_reads, K), np.int64)
    length[:, 0::2] = rng.integers(1, 2 * m_mean, size=(n_reads, n_gap + 1))
    length[:, 1::2] = rng.integers(1, 4, size=(n_reads, n_gap))
    ref_adv = np.where(code != OP_I, length, 0)
    qry_adv = np.where(code != OP_D, length, 0)
    span = ref_adv.sum(1)
    pos = np.sort(rng.integers(0, max(1, contig_len - int(span.max()) - 1), size=n_reads))
    qlen = qry_adv.sum(1)
    seq_off = np.concatenate([[0], np.cumsum(qlen)])
    a = Alignments()
    a.read_pos = pos.astype(np.int32)
    a.read_end = (pos + span).astype(np.int32)
    a.read_op0 = (np.arange(n_reads + 1) * K).astype(np.int32)
    a.read_strand = rng.integers(0, 2, size=n_reads).astype(np.uint8)
    a.op_ref = (pos[:, None] + np.cumsum(ref_adv, 1) - ref_adv).reshape(-1).astype(np.int32)
    a.op_qry = (seq_off[:-1, None] + np.cumsum(qry_adv, 1) - qry_adv).reshape(-1).astype(np.int32)
    a.op_len = ((length << 2) | code).reshape(-1).astype(np.int32)
    bases = np.frombuffer(b"ACGT", np.uint8)
    a.seq = bases[rng.integers(0, 4, size=int(seq_off[-1]))]
    reference = bases[rng.integers(0, 4, size=contig_len)].tobytes().decode()
    sites = np.arange(200, contig_len - 200, site_spacing) + rng.integers(0, site_spacing // 2, size=len(range(200, contig_len - 200, site_spacing)))
    return a, reference, np.unique(sites)


def alignments_to_sam(a, name="chr", reads=None):
    """SAM rows of (some reads of) an encoded region - how tests and the bench hand the same reads to the oracle."""
    from .create_tensor import OP_D
    rows = []
    for r in (range(a.n_reads) if reads is None else reads):
        lo, hi = int(a.read_op0[r]), int(a.read_op0[r + 1])
        cigar = "".join("%d%s" % (l >> 2, "MID"[l & 3]) for l in a.op_len[lo:hi].tolist())
        q0 = int(a.op_qry[lo])
        qn = int(sum(l >> 2 for l in a.op_len[lo:hi].tolist() if (l & 3) != OP_D))
        rows.append("\t".join(["r%d" % r, "16" if a.read_strand[r] else "0", name, str(int(a.read_pos[r]) + 1), "60", cigar,
                               "*", "0", "0", bytes(a.seq[q0:q0 + qn]).decode(), "*"]))
    return rows
