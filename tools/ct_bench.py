"""CreateTensor stage on the device: throughput of clairb_create_tensors on a synthetic ONT-like region,
its kernel time from the library's CUDA-event profile, the native SAM encoder, and - `cpu_part`, the only piece that touches
oracle/ - the pure-Python restatement timed on a bounded sample of the same region plus parity of that sample.
bench.py adds the result to its JSON line as `create_tensor_stage`.  Stand-alone: python tools/ct_bench.py [--contig N]"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


# read shapes of the three technologies the reference ships models for (README "Pretrained Models"); "ont" is the default
SHAPES = {"ont": dict(depth=40, read_len=8000, ops_per_read=1001, site_spacing=50),
          "ccs": dict(depth=30, read_len=14000, ops_per_read=61, site_spacing=300),
          "illumina": dict(depth=60, read_len=150, ops_per_read=3, site_spacing=500)}


def device_part(m, contig_len=1000000, repeats=5, with_forward=True, hbm_peak_gbs=None, shape="ont"):
    """-> (report, context for cpu_part).  `m` is a clair_b200.model.Clair (with weights when with_forward)."""
    from clair_b200 import create_tensor as CT, synth
    t0 = time.time()
    aln, reference, sites = synth.synthetic_alignments(contig_len, **SHAPES[shape])
    gen_s = time.time() - t0
    lib, h = m._lib, m._h
    n = int(sites.shape[0])
    if n > m.max_sites:
        raise ValueError("region holds %d sites, the model handle was sized for %d" % (n, m.max_sites))
    what = {"ont": "ONT-like region: %d bp, depth 40, 8 kb reads, an indel every ~8 bases, a candidate every ~50 bp",
            "ccs": "CCS-like region: %d bp, depth 30, 14 kb reads, an indel every ~230 bases, a candidate every ~300 bp",
            "illumina": "Illumina-like region: %d bp, depth 60, 150 bp reads, one indel per read, a candidate every ~500 bp"}[shape]
    out = {"workload": "synthetic " + what % contig_len,
           "sites": n, "reads": aln.n_reads, "ops": aln.n_ops, "query_bases": int(aln.seq.size), "synth_seconds": round(gen_s, 2)}
    in_bytes = sum(int(getattr(aln, f).nbytes) for f in aln.__slots__) + len(reference) + 4 * n
    out["input_bytes"] = in_bytes
    pinned = CT.Alignments()
    for f in aln.__slots__:
        src = getattr(aln, f)
        dst = CT._pinned_empty(src.shape[0], src.dtype)
        dst[...] = src
        setattr(pinned, f, dst)

    def timed(a, reps, **kw):
        ts = []
        for _ in range(reps):
            t = time.perf_counter()
            blk = CT.create_tensors(m, a, sites, reference, **kw)
            ts.append(time.perf_counter() - t)
        return float(np.median(ts)), blk

    block = CT.create_tensors(m, aln, sites, reference, subtract=True, fetch=False)      # warm-up: sizes the device buffers
    out["rows"] = len(block)
    lib.clairb_set_profiling(h, 1)
    call_s, _ = timed(pinned, repeats, subtract=True, fetch=False)
    buf = ctypes.create_string_buffer(1 << 16)
    lib.clairb_read_profile(h, buf, len(buf))
    lib.clairb_set_profiling(h, 0)
    k = {p["kernel"]: p for p in json.loads(buf.value.decode())}["create_tensors"]
    k_ms = k["ms"] / k["launches"]
    # algorithmic bytes of the kernel: every input byte once + one 2,112-byte row and one 8-byte meta record per site
    algo = in_bytes + n * (2112 + 8)
    out["kernel"] = {"name": "create_tensors", "ms": k_ms, "sites_per_s": n / (k_ms * 1e-3), "algorithmic_bytes": algo,
                     "achieved_gbs": algo / (k_ms * 1e-3) / 1e9, "bound": "hbm"}
    try:                                                     # DRAM bytes of one launch of this workload, from the committed ncu capture
        t = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic_ct_s19.json")))["create_tensors"]
        out["kernel"]["traffic"] = (t["dram_bytes_read"] + t["dram_bytes_write"]) if t["sites"] == n else None
        out["kernel"]["traffic_source"] = t["source"]
    except Exception:
        out["kernel"]["traffic"] = None
    if hbm_peak_gbs:
        out["kernel"]["hbm_peak_gbs"] = hbm_peak_gbs
        out["kernel"]["frac"] = out["kernel"]["achieved_gbs"] / hbm_peak_gbs
        # the same numbers under the key names of bench.py's `roofline` object
        out["roofline"] = {"bound": "hbm", "kernel": "create_tensors", "achieved": out["kernel"]["achieved_gbs"], "peak": hbm_peak_gbs,
                           "unit": "GB/s", "frac": out["kernel"]["frac"], "traffic": out["kernel"]["traffic"],
                           "algorithmic_bytes": algo, "avg_launch_ms": k_ms, "sites_per_launch": n,
                           "note": "bytes are at the algorithmic minimum; the kernel is bound by instruction issue under divergence "
                                   "(8.9 of 32 threads active per instruction, profiles/r01_s19_create_tensors.ncu_summary.txt)"}
    out["call_resident"] = {"seconds": call_s, "sites_per_s": n / call_s, "h2d_bytes": in_bytes,
                            "note": "clairb_create_tensors, encoded reads in pinned host memory (uploaded every call), tensors left on the device"}
    pg_s, _ = timed(aln, max(2, repeats // 2), subtract=True, fetch=False)
    out["call_resident_pageable"] = {"seconds": pg_s, "sites_per_s": n / pg_s, "note": "same, encoded reads in pageable numpy arrays"}
    ft_s, fetched = timed(pinned, max(2, repeats // 2), subtract=False, fetch=True)
    out["call_fetched"] = {"seconds": ft_s, "sites_per_s": n / ft_s, "d2h_bytes": n * 2112,
                           "note": "same, raw int16 rows copied back to the host (what CreateTensor.py prints as text)"}
    if with_forward:
        block = CT.create_tensors(m, pinned, sites, reference, subtract=True, fetch=False)
        keep = block.callable_sites()
        block.predict(keep)                                  # warm-up at full size: the gather / result buffers grow once
        t = time.perf_counter()
        block = CT.create_tensors(m, pinned, sites, reference, subtract=True, fetch=False)
        probs = block.predict(keep)
        dt = time.perf_counter() - t
        sub = CT.create_tensors(m, pinned, sites, reference, subtract=True, fetch=True)
        host_fed = m.predict_packed(sub.x[keep[:2000]])
        out["create_and_predict"] = {"seconds": dt, "sites_per_s": keep.shape[0] / dt, "sites": int(keep.shape[0]),
                                     "note": "encoded reads -> tensors -> forward -> probabilities on the host; the tensors never leave the device",
                                     "bit_identical_to_host_fed_forward": bool(np.array_equal(host_fed, probs[:2000]))}
    # the host encoder on the SAM text of a slice of the region (printing SAM rows from Python is slow, so only a slice)
    n_txt = min(aln.n_reads, 400)
    text = ("\n".join(synth.alignments_to_sam(aln, reads=range(n_txt))) + "\n").encode()
    enc = {}
    for encoder in ("native", "python"):
        t = time.perf_counter()
        e = CT.encode_alignments(text, encoder=encoder)
        dt = time.perf_counter() - t
        enc[encoder] = {"seconds": dt, "mb_per_s": len(text) / 1e6 / dt, "ops_per_s": e.n_ops / dt}
    enc["sample"] = "%d reads, %.1f MB of SAM text" % (n_txt, len(text) / 1e6)
    enc["note"] = "native = clairb_encode_sam (host threads, two passes); python = the numpy restatement it is checked against"
    out["encode_sam"] = enc
    return out, {"aln": aln, "reference": reference, "sites": sites, "fetched": fetched}


def cpu_part(out, ctx, cpu_sites=400):
    """The oracle (pure Python, like the reference) on `cpu_sites` sites from the middle of the region (full depth) and the
    reads that reach them; the device rows of those sites must be identical."""
    from clair_b200 import synth
    from oracle import create_tensor_oracle as O
    aln, reference, sites, fetched = ctx["aln"], ctx["reference"], ctx["sites"], ctx["fetched"]
    mid = sites.shape[0] // 2
    centres = sites[mid:mid + cpu_sites]
    reach = np.flatnonzero((aln.read_pos <= int(centres[-1]) + 40) & (aln.read_end >= int(centres[0]) - 40))
    sam = synth.alignments_to_sam(aln, reads=reach.tolist())
    t = time.perf_counter()
    want = O.create_tensors(sam, centres.tolist(), reference)
    cpu_s = time.perf_counter() - t
    at = np.searchsorted(fetched.positions, [r[1] for r in want])
    same = fetched.positions[at].tolist() == [r[1] for r in want] and \
        np.array_equal(fetched.x[at], np.stack([r[3] for r in want]).astype(np.int16))
    assert same, "device tensors differ from the CPU restatement of CreateTensor"
    out["cpu_oracle"] = {"sites": len(want), "reads_walked": int(reach.shape[0]), "seconds": cpu_s, "sites_per_s": len(want) / cpu_s,
                         "kind": "port (pure Python like the reference), 1 core", "device_rows_identical": bool(same)}
    return out


def hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return None


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--contig", type=int, default=1000000)
    ap.add_argument("--repeats", type=int, default=5)
    ap.add_argument("--no-forward", action="store_true")
    ap.add_argument("--shape", default="ont", choices=sorted(SHAPES))
    args = ap.parse_args()
    from clair_b200 import weights as W
    from clair_b200.model import Clair
    model = Clair(max_sites=max(4096, args.contig // 20), batch_sites=1000)
    if not args.no_forward:
        model.set_weights(W.random_weights(seed=1234))
    report, context = device_part(model, args.contig, args.repeats, not args.no_forward, hbm_peak(), args.shape)
    print(json.dumps(cpu_part(report, context)))
    model.close()
