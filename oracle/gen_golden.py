"""Generate the committed golden vectors under tests/golden/.  TEST INFRASTRUCTURE ONLY.

Run from the repo root:  python oracle/gen_golden.py
  forward_b8.npz   8 hand-shaped sites (zero depth, deep >=250, negative evidence, ragged noise)
                   + every layer output of the fp64 oracle for seed-1234 weights.
  decode_*.txt/npz text rows and what the REFERENCE's own clair/utils.py:tensor_generator_from
                   yields for them (imported from /root/reference with blosc/intervaltree stubbed;
                   only runs where /root/reference exists).
PARITY UNPINNED for the forward (no reference tests / TF): these vectors pin the oracle against
regressions and against the independent torch.nn.LSTM restatement in tests/test_oracle.py.
"""
import gzip
import hashlib
import io
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clair_b200 import synth, weights as W            # noqa: E402
from oracle import clair_oracle as O                  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"


def golden_inputs():
    counts = synth.synthetic_counts(8, seed=77).astype(np.int32)
    counts[0] = 0                                                  # zero depth everywhere
    counts[1] = np.minimum(counts[1] * 6, 32000)                   # deep pile-up (>=250)
    counts[1, :, :, 0] = np.maximum(counts[1, :, :, 0], 0)
    counts[2, 16, :, 1] = 0                                        # strong negative evidence at the centre
    counts[3, :, :, 3] += 40                                       # large positive SNP channel
    rng = np.random.default_rng(5)
    counts[4] = rng.integers(0, 300, size=counts[4].shape)         # unstructured
    return counts.astype(np.int16)


def weights_digest(w):
    h = hashlib.sha256()
    for k in sorted(w):
        h.update(k.encode())
        h.update(np.ascontiguousarray(w[k]).tobytes())
    return h.hexdigest()


def gen_forward():
    counts = golden_inputs()
    X = O.subtract_channel0(counts)
    w = W.random_weights(seed=1234)
    probs, im = O.forward(X, w, np.float64, intermediates=True)
    np.savez_compressed(
        os.path.join(GOLD, "forward_b8.npz"),
        counts=counts, X=X, weights_seed=np.int64(1234), weights_sha256=np.array(weights_digest(w)),
        lstm1=im["lstm1"].astype(np.float32), lstm2=im["lstm2"].astype(np.float32),
        l3=im["l3"].astype(np.float32), l4=im["l4"],
        logits=np.concatenate(im["logits"], axis=1), probs=np.concatenate(probs, axis=1))
    print("forward_b8.npz written; weights sha256", weights_digest(w)[:16])


def import_reference_utils():
    """Import the reference's clair/utils.py with its absent third-party imports stubbed."""
    sys.path.insert(0, REFERENCE)
    for name in ("blosc", "intervaltree"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.set_nthreads = lambda n: None
            m.IntervalTree = object
            sys.modules[name] = m
    import clair.utils as ref_utils
    return ref_utils


def gen_decode():
    if not os.path.isdir(REFERENCE):
        print("no /root/reference here: decode fixtures not regenerated")
        return
    ref_utils = import_reference_utils()
    counts = synth.synthetic_counts(11, seed=99)
    seqs = []
    rng = np.random.default_rng(3)
    for i in range(11):
        s = "".join(rng.choice(list("ACGT"), size=33))
        seqs.append(s)
    seqs[2] = seqs[2][:16] + "N" + seqs[2][17:]      # IUPAC, kept
    seqs[5] = seqs[5][:16] + "*" + seqs[5][17:]      # not IUPAC, dropped (utils.py:90)
    seqs[7] = seqs[7][:16] + "R" + seqs[7][17:]
    rows = []
    for i in range(11):
        rows.append("chr%d %d %s %s" % (i % 3 + 1, 1000 + 17 * i, seqs[i], " ".join("%d" % v for v in counts[i].reshape(-1))))
    text = "\n".join(rows) + "\n"
    path = os.path.join(GOLD, "decode_rows.txt.gz")
    with gzip.open(path, "wt") as f:
        f.write(text)
    out = {}
    err = io.StringIO()
    stderr, sys.stderr = sys.stderr, err
    try:
        for bi, (X, infos) in enumerate(ref_utils.tensor_generator_from(path, 4)):
            out["X%d" % bi] = np.array(X, copy=True)
            out["info%d" % bi] = np.array([" ".join(i) for i in infos])
    finally:
        sys.stderr = stderr
    out["stderr"] = np.array(err.getvalue())
    np.savez_compressed(os.path.join(GOLD, "decode_expected.npz"), **out)
    print("decode fixtures written:", sorted(out))


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    gen_forward()
    gen_decode()
