"""CreateTensor stage on the device: throughput of clairb_create_tensors on a synthetic ONT-like region, its kernel time
from the library's CUDA-event profile, the oracle (pure-Python restatement) timed on a bounded sample of the same region, and
parity of that sample.  Prints one JSON line.  Usage: python tools/ct_bench.py [--contig 1000000] [--repeats 5]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(contig_len=1000000, repeats=5, cpu_sites=400, with_forward=True, hbm_peak_gbs=None):
    import ctypes
    from clair_b200 import create_tensor as CT, synth, weights as W
    from clair_b200.model import Clair
    t0 = time.time()
    aln, reference, sites = synth.synthetic_alignments(contig_len)
    gen_s = time.time() - t0
    m = Clair(max_sites=max(4096, int(sites.shape[0])), batch_sites=1000)
    if with_forward:
        m.set_weights(W.random_weights(seed=1234))
    lib, h = m._lib, m._h
    out = {"workload": "synthetic ONT-like region: %d bp, depth 40, 8 kb reads, an indel every ~8 bases, a candidate every ~50 bp" % contig_len,
           "sites": int(sites.shape[0]), "reads": aln.n_reads, "ops": aln.n_ops, "query_bases": int(aln.seq.size),
           "synth_seconds": round(gen_s, 2)}
    in_bytes = sum(int(getattr(aln, f).nbytes) for f in aln.__slots__) + len(reference) + 4 * int(sites.shape[0])
    out["input_bytes"] = in_bytes
    # warm-up (allocates the growable buffers), then timed calls: device-resident result (fetch=False) and fetched
    block = CT.create_tensors(m, aln, sites, reference, subtract=True, fetch=False)
    out["rows"] = len(block)
    lib.clairb_set_profiling(h, 1)
    ts = []
    for _ in range(repeats):
        t = time.perf_counter()
        block = CT.create_tensors(m, aln, sites, reference, subtract=True, fetch=False)
        ts.append(time.perf_counter() - t)
    buf = ctypes.create_string_buffer(1 << 16)
    lib.clairb_read_profile(h, buf, len(buf))
    lib.clairb_set_profiling(h, 0)
    prof = {k["kernel"]: k for k in json.loads(buf.value.decode())}
    k = prof["create_tensors"]
    k_ms = k["ms"] / k["launches"]
    n = int(sites.shape[0])
    # algorithmic bytes of the kernel: every input byte once + one 2,112-byte row and one 8-byte meta record per site
    algo = in_bytes + n * (2112 + 8)
    out["kernel"] = {"name": "create_tensors", "ms": k_ms, "sites_per_s": n / (k_ms * 1e-3), "algorithmic_bytes": algo,
                     "achieved_gbs": algo / (k_ms * 1e-3) / 1e9}
    if hbm_peak_gbs:
        out["kernel"]["hbm_peak_gbs"] = hbm_peak_gbs
        out["kernel"]["frac"] = out["kernel"]["achieved_gbs"] / hbm_peak_gbs
    call_s = float(np.median(ts))
    out["call_resident"] = {"seconds": call_s, "sites_per_s": n / call_s,
                            "note": "clairb_create_tensors with host inputs (reads uploaded every call), tensors left on the device"}
    ts = []
    for _ in range(max(2, repeats // 2)):
        t = time.perf_counter()
        fetched = CT.create_tensors(m, aln, sites, reference, subtract=False, fetch=True)
        ts.append(time.perf_counter() - t)
    out["call_fetched"] = {"seconds": float(np.median(ts)), "sites_per_s": n / float(np.median(ts)),
                           "d2h_bytes": n * 2112, "note": "same, int16 rows copied back to the host"}
    if with_forward:
        block = CT.create_tensors(m, aln, sites, reference, subtract=True, fetch=False)
        keep = block.callable_sites()
        block.predict(keep[:1024])
        t = time.perf_counter()
        block = CT.create_tensors(m, aln, sites, reference, subtract=True, fetch=False)
        probs = block.predict(keep)
        dt = time.perf_counter() - t
        out["create_and_predict"] = {"seconds": dt, "sites_per_s": keep.shape[0] / dt, "sites": int(keep.shape[0]),
                                     "note": "alignments -> tensors -> forward, tensors never leave the device"}
        sub = CT.create_tensors(m, aln, sites, reference, subtract=True, fetch=True)
        host_fed = m.predict_packed(sub.x[keep[:2000]])
        out["create_and_predict"]["bit_identical_to_host_fed_forward"] = bool(np.array_equal(host_fed, probs[:2000]))
    # CPU: the oracle on `cpu_sites` sites from the middle of the region (full depth) and the reads that reach them
    from oracle import create_tensor_oracle as O
    mid = n // 2
    cpu_centres = sites[mid:mid + cpu_sites]
    reach = np.flatnonzero((aln.read_pos <= int(cpu_centres[-1]) + 40) & (aln.read_end >= int(cpu_centres[0]) - 40))
    n_reads = int(reach.shape[0])
    sam = synth.alignments_to_sam(aln, reads=reach.tolist())
    t = time.perf_counter()
    want = O.create_tensors(sam, cpu_centres.tolist(), reference)
    cpu_s = time.perf_counter() - t
    at = np.searchsorted(fetched.positions, [r[1] for r in want])
    same = fetched.positions[at].tolist() == [r[1] for r in want] and np.array_equal(fetched.x[at], np.stack([r[3] for r in want]).astype(np.int16))
    out["cpu_oracle"] = {"sites": len(want), "reads_walked": n_reads, "seconds": cpu_s, "sites_per_s": len(want) / cpu_s,
                         "kind": "port (pure Python like the reference), 1 core", "device_rows_identical": bool(same)}
    m.close()
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--contig", type=int, default=1000000)
    ap.add_argument("--repeats", type=int, default=5)
    ap.add_argument("--no-forward", action="store_true")
    args = ap.parse_args()
    peak = None
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    print(json.dumps(run(args.contig, args.repeats, with_forward=not args.no_forward, hbm_peak_gbs=peak)))
