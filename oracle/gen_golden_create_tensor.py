"""Golden vectors for CreateTensor, produced by the REFERENCE's own code.  TEST INFRASTRUCTURE.

Run from the repo root (only where /root/reference exists):  python oracle/gen_golden_create_tensor.py
Imports /root/reference/dataPrepScripts/CreateTensor.py unmodified and runs its OutputAlnTensor(args) on synthetic
alignments.  The three child processes it spawns (`samtools faidx`, `samtools view`, `gzip -fdc candidates`) are replaced
by in-memory line sources through the module's `subprocess_popen` name; everything after the pipes - candidate
bookkeeping, CIGAR walk, depth cap, generate_tensor, row formatting - is the reference's code.  The rows it prints are
stored in tests/golden/create_tensor_cases.json.gz next to the inputs that produced them.
"""
import gzip
import io
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"


def import_reference_create_tensor():
    sys.path.insert(0, REFERENCE)
    import dataPrepScripts.CreateTensor as CT
    return CT


class _Lines(object):
    def __init__(self, lines):
        self._it = iter(lines)

    def __iter__(self):
        return self._it

    def close(self):
        pass


class _Proc(object):
    returncode = 0

    def __init__(self, lines):
        self.stdout = _Lines(lines)

    def wait(self):
        return 0


class _Sink(io.StringIO):
    def close(self):            # TensorStdout.__del__ closes its handle; keep the text readable
        pass


def run_reference(CT, case):
    """One OutputAlnTensor run of the reference on an in-memory case; returns the rows it wrote."""
    contig = case["contig"]

    def popen(args, **kw):
        if "faidx" in args:
            region = args[-1]
            lo, hi = 1, len(contig)
            if ":" in region:
                lo, hi = (int(v) for v in region.split(":")[1].split("-"))
            seq = contig[lo - 1:hi]
            return _Proc([">%s\n" % region] + [seq[i:i + 60] + "\n" for i in range(0, len(seq), 60)])
        if "view" in args:
            return _Proc([l + "\n" for l in case["sam"]])
        if args[0] == "gzip":
            return _Proc([l + "\n" for l in case["candidates"]])
        raise AssertionError("unexpected child process %r" % (args,))

    a = case["args"]
    ns = types.SimpleNamespace(samtools="samtools", tensor_fn="PIPE", bam_fn="x.bam", ref_fn="x.fa", can_fn="x.can",
                               dcov=a["dcov"], stop_consider_left_edge=a["stop_consider_left_edge"],
                               minCoverage=a["minCoverage"], minMQ=a["minMQ"], ctgName=a["ctgName"],
                               ctgStart=a["ctgStart"], ctgEnd=a["ctgEnd"])
    saved = (CT.subprocess_popen, CT.param.expandReferenceRegion, sys.stdout)
    sink = _Sink()
    CT.subprocess_popen = popen
    CT.param.expandReferenceRegion = a["expandReferenceRegion"]
    sys.stdout = sink
    try:
        CT.OutputAlnTensor(ns)
    finally:
        CT.subprocess_popen, CT.param.expandReferenceRegion, sys.stdout = saved
    return sink.getvalue().splitlines()


def random_contig(rng, length):
    s = rng.choice(list("ACGT"), size=length)
    for ch, rate in (("N", 0.01), ("a", 0.01), ("g", 0.01), ("M", 0.003), ("R", 0.003), ("X", 0.003)):
        s[rng.random(length) < rate] = ch
    return "".join(s)


def random_read(rng, contig_len, pos0, style):
    """(cigar, seq) of a read starting at 0-based pos0 that stays inside the contig."""
    room = contig_len - pos0
    ops = []
    if rng.random() < 0.15:
        ops.append((int(rng.integers(1, 20)), "H"))
    if rng.random() < 0.3:
        ops.append((int(rng.integers(1, 12)), "S"))
    if style == "leading" and rng.random() < 0.5:
        ops.append((int(rng.integers(1, 4)), "I" if rng.random() < 0.5 else "D"))
    n_ops = int(rng.integers(1, 40 if style != "short" else 6))
    used = sum(n for n, c in ops if c == "D")
    for _ in range(n_ops):
        r = rng.random()
        if r < 0.55:
            c = "M" if rng.random() < 0.8 else ("=" if rng.random() < 0.5 else "X")
            n = int(rng.integers(1, 45))
        elif r < 0.75:
            c, n = "I", int(rng.integers(1, 6)) if rng.random() < (0.5 if style == "long_indels" else 0.9) else int(rng.integers(20, 130))
        elif r < 0.95:
            c, n = "D", int(rng.integers(1, 6)) if rng.random() < (0.5 if style == "long_indels" else 0.9) else int(rng.integers(20, 130))
        elif r < 0.98:
            c, n = "N", int(rng.integers(1, 30))
        else:
            c, n = "P", int(rng.integers(1, 3))
        if c in "M=XD":
            if used + n > room:
                n = room - used
                if n <= 0:
                    break
            used += n
        ops.append((n, c))
    if rng.random() < 0.3:
        ops.append((int(rng.integers(1, 12)), "S"))
    qlen = sum(n for n, c in ops if c in "M=XIS")
    if qlen == 0:
        ops.append((1, "S"))
        qlen = 1
    seq = rng.choice(list("ACGT"), size=qlen)
    for ch, rate in (("N", 0.01), ("c", 0.02), ("t", 0.02), ("Y", 0.004), (".", 0.003), ("=", 0.003)):
        seq[rng.random(qlen) < rate] = ch
    cigar = "".join("%d%s" % (n, c) for n, c in ops)
    if style == "odd":
        # tokens the SAM specification does not produce; the reference's character loop (:283-366) gives them no effect
        junk = ["3B", "0M", "2m", "7i", "*", "0D", "12d", "4Z", "0I"]
        tokens = ["%d%s" % (n, c) for n, c in ops]
        for _ in range(int(rng.integers(1, 5))):
            tokens.insert(int(rng.integers(0, len(tokens) + 1)), junk[int(rng.integers(0, len(junk)))])
        cigar = "".join(tokens)
    return cigar, "".join(seq)


def make_case(rng, name, contig_len, n_reads, n_cand, style="mixed", dcov=250, min_mq=0, min_cov=0,
              region=None, expand=1000000, stop_left=False, dup_rate=0.0):
    contig = random_contig(rng, contig_len)
    starts = np.sort(rng.integers(0, contig_len - 5, size=n_reads))
    if style == "from_zero":
        starts[:3] = 0
    sam = ["@SQ\tSN:%s\tLN:%d" % (name, contig_len)]
    prev = None
    for i, p in enumerate(starts):
        if prev is not None and rng.random() < dup_rate:
            p = prev
        prev = int(p)
        cigar, seq = random_read(rng, contig_len, int(p), style)
        flag = int(rng.choice([0, 16, 0, 16, 1, 83, 99, 1024 + 16]))
        mq = int(rng.integers(0, 61))
        sam.append("\t".join(["r%d" % i, str(flag), name, str(int(p) + 1), str(mq), cigar, "*", "0", "0", seq, "*"]))
    lo, hi = (1, contig_len) if region is None else region
    cpos = np.unique(rng.integers(max(1, lo - 30), min(contig_len, hi + 30) + 1, size=n_cand))
    if style == "dense":
        base = int(rng.integers(lo + 40, hi - 80))
        cpos = np.unique(np.concatenate([cpos, np.arange(base, base + 40)]))
    cands = ["%s\t%d\tA\t3\t10" % (name, p) for p in cpos]
    if n_cand > 4:
        cands.insert(3, cands[2])                  # a duplicated candidate row
    return {"name": name, "contig": contig, "sam": sam, "candidates": cands,
            "args": {"dcov": dcov, "stop_consider_left_edge": stop_left, "minCoverage": min_cov, "minMQ": min_mq,
                     "ctgName": name, "ctgStart": None if region is None else region[0],
                     "ctgEnd": None if region is None else region[1], "expandReferenceRegion": expand}}


def make_cases():
    rng = np.random.default_rng(20240611)
    return [
        make_case(rng, "whole", 1500, 160, 60),
        make_case(rng, "dense", 1200, 200, 30, style="dense"),
        make_case(rng, "short_reads", 900, 300, 50, style="short"),
        make_case(rng, "leading_indels", 900, 150, 60, style="leading"),
        make_case(rng, "from_zero", 600, 80, 40, style="from_zero", dup_rate=0.3, dcov=3),
        make_case(rng, "depth_cap", 800, 260, 40, dup_rate=0.6, dcov=4),
        make_case(rng, "min_mq_cov", 1000, 220, 60, min_mq=20, min_cov=6),
        make_case(rng, "region", 2500, 260, 70, region=(600, 1900)),
        make_case(rng, "region_tight_ref", 2500, 260, 70, region=(700, 1800), expand=40, style="short"),
        make_case(rng, "no_left_edge", 1200, 200, 60, stop_left=True),
        make_case(rng, "no_reads", 400, 0, 10),
        # second batch (appended: the cases above keep their random streams)
        make_case(rng, "long_indels", 2000, 260, 120, style="long_indels"),
        make_case(rng, "deep", 260, 600, 40, style="short"),
        make_case(rng, "deep_long", 700, 500, 60),
        make_case(rng, "contig_tail", 300, 120, 200, style="dense"),
        make_case(rng, "dcov_one", 700, 200, 50, dup_rate=0.5, dcov=1),
        make_case(rng, "region_from_one", 1500, 200, 80, region=(1, 700), expand=25),
        make_case(rng, "no_left_edge_dense", 900, 260, 40, style="dense", stop_left=True, dup_rate=0.2, dcov=3),
        make_case(rng, "odd_cigars", 900, 220, 60, style="odd"),
    ]


def main():
    if not os.path.isdir(REFERENCE):
        print("no /root/reference here: create_tensor fixtures not regenerated")
        return
    CT = import_reference_create_tensor()
    cases = make_cases()
    rows = 0
    for case in cases:
        case["expected"] = run_reference(CT, case)
        rows += len(case["expected"])
        print("%-18s reads %4d candidates %3d -> %3d rows" % (case["name"], len(case["sam"]) - 1, len(case["candidates"]),
                                                             len(case["expected"])))
    os.makedirs(GOLD, exist_ok=True)
    path = os.path.join(GOLD, "create_tensor_cases.json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(json.dumps(cases).encode())
    print("%s written: %d cases, %d rows, %d bytes" % (path, len(cases), rows, os.path.getsize(path)))


if __name__ == "__main__":
    main()
