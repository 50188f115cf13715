#!/usr/bin/env python
"""bench.py - candidate-sites/sec of the Clair forward path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): ONT-shape model, 1M synthetic candidate sites in predict-batches of 1000.  A "step"
is one pass of the forward over all `--batches-per-step` predict-batches (default 1000 x 1000 sites = 4.2 GB of fp32
input, far larger than the 126 MB L2; the sites are `--distinct-batches` distinct synthetic batches repeated); the default
20 steps keep the timed region above 2 s, so the sustained tensor peak is the right roofline denominator.
  value    : device-resident sites/s (inputs already in HBM, CUDA events on the launching stream)
  e2e      : one Clair.predict call per step over the whole pool (four fresh float32 arrays per call like
             clair/model.py:946-966): pinned HOST input in the int16 transport (the same integer counts, bit-identical
             results), H2D + forward + D2H inside the timed region; `e2e_f32` is the same call fed float32
  loop_e2e : the reference-shaped loop - a generator yielding 1000-site batches, one predict per batch
             (clair/call_var.py:1340-1344, param.predictBatchSize), a no-op output stage - through
             clair_b200.call_var.run_batches (predict_async: many batches in flight); `loop_lockstep` is the same loop
             with exactly one predict in flight, the reference's own structure
  roofline / cpu_baseline / oracle_pinning : see DESIGN.md section "Measurement"
N>1 (launched by torchrun, one rank per GPU): every rank runs the same per-GPU workload on its own sites (weak scaling);
the packed [sites,90] probabilities are gathered to rank 0 with one NCCL collective per step, inside the timed region
(e2e: host input -> device -> gather -> rank 0's host memory), and rank 0 checks the gathered rows bit for bit against
a single-GPU run of the same sites.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLOP_PER_SITE = 40386432            # 2 x 20,193,216 MAC
BYTES_PER_SITE = 4584               # 4,224 in + 360 out
LSTM2_FLOP_PER_SITE = 2 * 12976128  # SURVEY.md 8a row a7 (dominant kernel)
LSTM1_FLOP_PER_SITE = 2 * 5406720
KERNEL_FLOP_PER_SITE = {"lstm_layer2": LSTM2_FLOP_PER_SITE, "lstm_layer1": LSTM1_FLOP_PER_SITE,
                        "l4_dense": 2 * 1474560, "l3_slice_dense": 2 * 253440, "tail_heads": 2 * (73728 + 8640),
                        "lstm_seq1": 2 * 5406720, "xproj2": 2 * 8650752, "lstm_seq2": 2 * 4325376,
                        "l3l4_fused": 2 * (253440 + 1474560), "prep_tiles": 0, "prep_input": 0, "heads_tc": 2 * (73728 + 8640),
                        "lstm_seq_x2": LSTM2_FLOP_PER_SITE}
# algorithmic HBM bytes per site of each kernel (operand tiles in + results out)
KERNEL_BYTES_PER_SITE = {"prep_tiles": 4224 + 6336, "lstm_seq1": 6336 + 33792, "xproj2": 33792 + 135168,
                         "lstm_seq2": 135168 + 40960, "l3l4_fused": 40960 + 768 + 768, "heads_tc": 768 + 360 + 360,
                         "lstm_seq_x2": 2 * 33792 + 40960}      # h1 read once per direction, h2 tiles written
BATCH = 1000                        # shared/param.py:16 predictBatchSize
WORKLOAD = "ONT-shape model inference, 1M synthetic candidate sites, batch=1000, %dxB200"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batches-per-step", type=int, default=1000)
    ap.add_argument("--distinct-batches", type=int, default=200, help="distinct synthetic predict-batches, repeated to fill a step")
    ap.add_argument("--in-flight", type=int, default=64, help="predict-batches in flight in the loop_e2e leg")
    ap.add_argument("--cpu-baseline-batches", type=int, default=0, help="0 = auto (about 10-20 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-create-tensor", action="store_true", help="skip the CreateTensor-stage line (SURVEY 8f row 4)")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step line (SURVEY 8f row 5)")
    return ap.parse_args()


TRAFFIC_FILES = ("r02_traffic.json", "r01_traffic_s11.json")     # newest capture first


def measured_traffic(kernel, sites_per_launch):
    """DRAM bytes per launch of `kernel` from the newest committed ncu --set full capture that covers it (profiles/),
    scaled to this run's sites per launch; (None, None) when no capture covers the kernel."""
    for name in TRAFFIC_FILES:
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                t = json.load(f)
            return t["dram_bytes_per_site"][kernel] * sites_per_launch, name
        except (OSError, KeyError, ValueError):
            continue
    return None, None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p.get("hbm_gbs", 6650.0), "bf16_burst": p.get("bf16_tflops", 1590.0),
                "bf16_sustained": p.get("bf16_tflops_sustained", 1400.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback"}


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md).  The sampler is started
    early (nvidia-smi needs a few hundred ms to come up, the timed region is ~130 ms) and every sample carries a
    timestamp; stop() keeps the samples that fall between mark_begin() and mark_end()."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.t_begin = self.t_end = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    @staticmethod
    def _stamp(text, fallback):
        import datetime
        try:
            return datetime.datetime.strptime(text.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return fallback

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for seen, line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 10:
                continue
            try:
                rows.append((self._stamp(parts[0], seen), float(parts[2]), float(parts[3]), float(parts[4]), parts[6:10]))
            except ValueError:
                continue
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        lo, hi = self.t_begin or 0.0, self.t_end or float("inf")
        inside = [r for r in rows if lo - 0.005 <= r[0] <= hi + 0.005]
        window = "timed region"
        if not inside:
            # no sample landed inside: take the samples closest to it (the run before / after is the same workload)
            mid = 0.5 * (lo + hi)
            inside = sorted(rows, key=lambda r: abs(r[0] - mid))[:3]
            window = "nearest to the timed region"
        sm = sorted(r[1] for r in inside)
        reasons = set()
        for r in inside:
            for name, val in zip(names, r[4]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(r[2] for r in inside), "power_w_max": max(r[3] for r in inside),
                "samples": len(inside), "window": window, "reasons": sorted(reasons)}


def cpu_info():
    model = "unknown"
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return model, os.cpu_count() or 1


# ---- the reference's CPU path on the host cores ---------------------------------------------------------------------
# The reference scales on a host by running independent processes over genome chunks (`parallel -j`,
# /root/reference README "Run Clair ... in parallel", callVarBamParallel.py:90-119), each with a few intra-op threads
# (--threads, 4 in the README's commands).  The arm mirrors that: cores // 4 worker processes x 4 threads, every worker
# running the fp32 CPU port of the forward (oracle/clair_oracle_fast.py; TensorFlow 1.13.2 itself is not installable) on its
# own predict-batches.
def _cpu_worker(rank, threads, weights, seeds, check_batch, conn):
    """seeds: the synthetic predict-batches of this worker (clair_b200.synth.synthetic_tensors(BATCH, seed))"""
    import numpy as np                                     # noqa: F401
    import torch
    torch.set_num_threads(threads)
    sys.path.insert(0, ROOT)
    from clair_b200 import synth
    from oracle.clair_oracle_fast import FastOracle
    fo = FastOracle(weights, threads=threads)
    pool = [synth.synthetic_tensors(BATCH, seed=s) for s in seeds]
    if check_batch is not None:
        pool[0] = check_batch
    first = fo.forward_packed(pool[0])                     # warm-up, and the checker's sample
    conn.send(("ready", first if rank == 0 else None))
    k = 0
    while True:
        msg = conn.recv()
        if msg is None:
            return
        keep = isinstance(msg, tuple)                    # ("keep", n): also hand back the results (parity sample)
        n_batches = msg[1] if keep else msg
        outs = []
        t0 = time.perf_counter()
        for _ in range(n_batches):
            o = fo.forward_packed(pool[k % len(pool)])
            if keep:
                outs.append(o)
            k += 1
        conn.send(("done", time.perf_counter() - t0, outs))


class CpuArm(object):
    """cores // 4 processes x 4 threads of the CPU port; run(k) = every worker does k predict-batches, returns seconds."""

    def __init__(self, weights, seeds, procs=None, threads=4, check_batch=None):
        import multiprocessing as mp
        cores = os.cpu_count() or 1
        self.threads = min(threads, cores)
        self.procs = procs or max(1, cores // self.threads)
        ctx = mp.get_context("spawn")
        self.workers = []
        for r in range(self.procs):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_cpu_worker, daemon=True,
                            args=(r, self.threads, weights, [s + 1000 * r for s in seeds], check_batch if r == 0 else None, b))
            p.start()
            self.workers.append((p, a))
        self.first = None
        for r, (_, a) in enumerate(self.workers):
            tag, first = a.recv()
            if r == 0:
                self.first = first

    def run(self, k, keep=False):
        """every worker does k predict-batches; seconds (and, with keep, the [k*BATCH,90] results of every worker)"""
        t0 = time.perf_counter()
        for _, a in self.workers:
            a.send(("keep", k) if keep else k)
        outs = [a.recv()[2] for _, a in self.workers]
        dt = time.perf_counter() - t0
        return (dt, outs) if keep else dt

    def close(self):
        for p, a in self.workers:
            try:
                a.send(None)
            except OSError:
                pass
        for p, _ in self.workers:
            p.join(timeout=10)


def run_reference(args, rank):
    """--impl reference: the reference's CPU path on all host cores (see CpuArm), a bounded sample per step."""
    if rank != 0:
        return
    from clair_b200 import weights as W
    w = W.random_weights(seed=1234)
    arm = CpuArm(w, [20240607 + i for i in range(2)])
    per_worker = 1                                      # predict-batches per worker per step
    try:
        for _ in range(max(1, args.warmup)):
            arm.run(per_worker)
        dt = 0.0
        for _ in range(args.steps):
            dt += arm.run(per_worker)
    finally:
        arm.close()
    per_step = per_worker * arm.procs
    value = args.steps * per_step * BATCH / dt
    model, cores = cpu_info()
    sample = "%d steps x %d predict-batches x %d sites (of the 1M-site workload); %d processes x %d threads" % (
        args.steps, per_step, BATCH, arm.procs, arm.threads)
    print(json.dumps({
        "impl": "reference", "metric": "candidate-sites/sec", "value": value, "unit": "sites/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD % 1, "batch": BATCH, "sites_per_step": per_step * BATCH,
                   "note": "reference CPU path = fp32 torch-CPU port of the reference graph (TensorFlow 1.13.2 not installable), "
                           "run the way the reference scales on a host: independent processes x 4 threads (README `parallel -j`)"},
        "cpu_baseline": {"value": value, "unit": "sites/s", "cores": arm.procs * arm.threads, "kind": "port", "sample": sample,
                         "cpu_model": model, "host_cores": cores},
        "e2e": {"value": value, "unit": "sites/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def oracle_pinning(m, weights, X256):
    """How far the forward oracle can be pinned on this box (the reference's arithmetic lives in TensorFlow
    1.13): (1) is TensorFlow importable (then the reference itself could be run); (2) cuDNN's own fp32 LSTM + torch dense
    trunk on this GPU as a third, independent restatement against the fp64 numpy oracle and against the product."""
    import numpy as np
    out = {}
    try:
        import tensorflow as tf                                   # noqa: F401
        out["tensorflow"] = getattr(tf, "__version__", "present")
    except Exception as exc:
        out["tensorflow"] = "not importable (%s)" % type(exc).__name__
    try:
        from oracle import clair_oracle as O
        from oracle.clair_oracle_cudnn import CudnnOracle
        from clair_b200 import _lib
        ref_probs, im = O.forward(X256, weights, np.float64, intermediates=True)
        ref_logits = np.concatenate(im["logits"], axis=1)
        cp, cl = CudnnOracle(weights, device="cuda").forward(X256)
        m.predict(X256)
        ours = m.get_layer(_lib.LAYER_LOGITS, X256.shape[0])
        scaled = lambda a, b: float((np.abs(a - b) / np.maximum(1.0, np.abs(b))).max())
        out.update({"sites": int(X256.shape[0]),
                    "cudnn_fp32_vs_fp64_oracle_logits": scaled(cl, ref_logits),
                    "ours_vs_fp64_oracle_logits": scaled(ours, ref_logits),
                    "ours_vs_cudnn_fp32_logits": scaled(ours, cl),
                    "argmax_identical_all_three": bool(all(
                        (cp[:, a:b].argmax(1) == np.concatenate(ref_probs, axis=1)[:, a:b].argmax(1)).all()
                        for a, b in ((0, 21), (21, 24), (24, 57), (57, 90)))),
                    "note": "torch.nn.LSTM on cuda = cuDNN's LSTM, the kernel CudnnCompatibleLSTMCell (clair/model.py:282-312) is "
                            "defined to be weight-compatible with; TF32 off"})
    except Exception as exc:
        out["cudnn_error"] = "%s: %s" % (type(exc).__name__, exc)
    return out


def _decision_codes(infos):
    from clair_b200 import decision
    return decision.ref_base_codes(infos)


def main():
    args = parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # direct invocation with --gpus N: relaunch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29400 + os.getpid() % 500), os.path.abspath(__file__)]
        cmd += sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from clair_b200 import _lib, call_var, synth, weights as W
    from clair_b200.model import Clair, pinned_empty

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(seconds):
        if world == 1:
            return seconds
        t = torch.tensor([seconds], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    bps = args.batches_per_step
    sites = bps * BATCH
    steps, warm = args.steps, max(args.warmup, 3)
    weights = W.random_weights(seed=1234)
    m = Clair(device=local_rank, max_sites=sites, batch_sites=BATCH)
    m.set_weights(weights)
    lib, h = m._lib, m._h

    # ---- synthetic pool: `distinct` predict-batches per rank (different on every rank), repeated to fill the step ----
    distinct = min(args.distinct_batches, bps)
    Xi = pinned_empty((sites, 33, 8, 4), np.int16)                 # the int16 transport of the same counts
    base = synth.synthetic_counts(distinct * BATCH, seed=20240607 + 1 + 1000 * rank)
    base[..., 1:] -= base[..., 0:1]                                # clair/utils.py:96-98
    for s0 in range(0, sites, distinct * BATCH):
        k = min(distinct * BATCH, sites - s0)
        Xi[s0:s0 + k] = base[:k]
    del base
    X = pinned_empty((sites, 33, 8, 4), np.float32)                # what the reference's generator yields
    X[...] = Xi
    xd = torch.from_numpy(X).cuda(non_blocking=False)
    od = torch.empty((sites, _lib.N_OUT), dtype=torch.float32, device="cuda")
    gather_bufs = [torch.empty_like(od) for _ in range(world)] if (world > 1 and rank == 0) else None
    stream = torch.cuda.current_stream()

    def device_step():
        m.predict_device(xd.data_ptr(), _lib.DTYPE_F32, sites, od.data_ptr(), stream.cuda_stream)
        if world > 1:
            dist.gather(od, gather_list=gather_bufs, dst=0)     # the path's one exchange

    sampler = ClockSampler(local_rank)       # started well before the timed region: nvidia-smi takes a while to come up
    if rank == 0:
        sampler.start()

    # ---- device-resident timing ----
    for _ in range(warm):
        device_step()
    torch.cuda.synchronize()
    barrier()
    m.set_profiling(True)
    launches0 = m.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    sampler.mark_begin()
    ev0.record(stream)
    for _ in range(steps):
        device_step()
    ev1.record(stream)
    torch.cuda.synchronize()
    sampler.mark_end()
    barrier()
    launches = m.kernel_launches() - launches0
    ms = ev0.elapsed_time(ev1)
    profile = m.read_profile()
    m.set_profiling(False)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * sites * steps / (ms * 1e-3)

    # ---- end to end: host input -> probabilities in host memory, once per step over the whole pool ----
    # N = 1: Clair.predict (four fresh arrays).  N > 1: every rank runs host input -> device rows (clairb_predict_to_device),
    # ONE NCCL gather of the packed rows to rank 0, and rank 0 copies all N x [sites,90] to its pinned host memory.
    host_all = pinned_empty((world, sites, _lib.N_OUT), np.float32) if (world > 1 and rank == 0) else None
    host_all_t = torch.from_numpy(host_all) if host_all is not None else None
    last = {}

    d2h_stream = torch.cuda.Stream() if world > 1 else None
    E2E_SLICES = 16     # N > 1: gather + rank 0's device->host copy of slice j run while slice j+1 is computed

    def e2e_step(Xh, dtype):
        if world == 1:
            last["out4"] = m.predict(Xh)
            return
        per = -(-sites // E2E_SLICES)
        bounds = [(lo, min(sites, lo + per)) for lo in range(0, sites, per)]
        tickets = []
        for lo, hi in bounds:                          # every slice is queued at once: the pipeline of the handle never drains
            t = ctypes.c_int64()
            _lib.check(lib.clairb_predict_async_to_device(h, Xh[lo:hi].ctypes.data, dtype, hi - lo, od[lo:hi].data_ptr(),
                                                          ctypes.byref(t)), h, "clairb_predict_async_to_device")
            tickets.append(t.value)
        for (lo, hi), t in zip(bounds, tickets):
            _lib.check(lib.clairb_predict_wait(h, t), h, "clairb_predict_wait")
            dist.gather(od[lo:hi], gather_list=[b[lo:hi] for b in gather_bufs] if rank == 0 else None, dst=0)
            if rank == 0:
                # rank 0's copies to the host go out on their own stream: the next slice's gather must not queue behind them
                ev = torch.cuda.Event()
                ev.record(stream)
                d2h_stream.wait_event(ev)
                with torch.cuda.stream(d2h_stream):
                    for r in range(world):
                        host_all_t[r, lo:hi].copy_(gather_bufs[r][lo:hi], non_blocking=True)
        torch.cuda.synchronize()

    def time_e2e(Xh, dtype, n_steps):
        for _ in range(2):
            e2e_step(Xh, dtype)
        torch.cuda.synchronize()
        barrier()
        l0 = m.kernel_launches()
        t0 = time.perf_counter()
        for _ in range(n_steps):
            e2e_step(Xh, dtype)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        return world * sites * n_steps / dt, m.kernel_launches() - l0

    e2e_value, e2e_launches = time_e2e(Xi, _lib.DTYPE_I16, steps)
    out = np.concatenate(last["out4"], axis=1) if world == 1 else None
    e2e_f32_value, _ = time_e2e(X, _lib.DTYPE_F32, steps)
    if world == 1:
        assert np.array_equal(np.concatenate(last["out4"], axis=1), out), "int16 transport must give bit-identical results"
        assert np.isfinite(out).all()

    # ---- multi-GPU parity gate (BASELINE.md 3 / SURVEY.md 8d): the gathered rows of every rank must equal, bit for bit,
    #      what ONE GPU computes for the same sites.  Rank 0 collects every rank's input and runs it on its own device. ----
    multi_gpu_check = None
    if world > 1:
        xi_d = torch.from_numpy(Xi).cuda().view(torch.uint8)          # NCCL has no int16: the bytes travel
        in_bufs = [torch.empty_like(xi_d) for _ in range(world)] if rank == 0 else None
        dist.gather(xi_d, gather_list=in_bufs, dst=0)
        if rank == 0:
            single = torch.empty_like(od)
            bad = 0
            for r in range(world):
                m.predict_device(in_bufs[r].data_ptr(), _lib.DTYPE_I16, sites, single.data_ptr(), stream.cuda_stream)
                torch.cuda.synchronize()
                if not torch.equal(single, gather_bufs[r]) or not np.array_equal(host_all[r], single.cpu().numpy()):
                    bad += 1
            multi_gpu_check = {"rows_checked": world * sites, "ranks_differing": bad,
                               "what": "gathered [N*sites,90] of the e2e leg (device buffers and rank 0's host copy) vs a "
                                       "single-GPU run of every rank's sites on rank 0, bit for bit"}
            assert bad == 0, "multi-GPU result differs from the single-GPU result"
            del in_bufs, single
        del xi_d

    # ---- what the host link allows: pinned H2D bandwidth of this box ----
    xt = torch.from_numpy(X)
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    xd.copy_(xt, non_blocking=True)
    torch.cuda.synchronize()
    barrier()
    h0.record(stream)
    for _ in range(3):
        xd.copy_(xt, non_blocking=True)
    h1.record(stream)
    torch.cuda.synchronize()
    h2d_gbs = 3 * sites * 4224 / (h0.elapsed_time(h1) * 1e-3) / 1e9

    # ---- the reference-shaped loop: generator -> one predict per 1000-site batch -> output stage (no-op) ----
    def batch_source(Xh, n_batches, pageable=False):
        infos = [None] * BATCH                                   # the forward never looks at them
        for i in range(n_batches):
            j = i % bps
            xb = Xh[j * BATCH:(j + 1) * BATCH]
            yield (np.array(xb) if pageable else xb), infos

    def time_loop(Xh, n_batches, in_flight, pageable=False):
        seen = [0]

        def output_stage(batch, prediction):
            seen[0] += prediction[0].shape[0]

        call_var.run_batches(m, batch_source(Xh, 2 * args.in_flight, pageable), output_stage, in_flight=in_flight)   # warm-up
        seen[0] = 0
        torch.cuda.synchronize()
        barrier()
        l0 = m.kernel_launches()
        t0 = time.perf_counter()
        call_var.run_batches(m, batch_source(Xh, n_batches, pageable), output_stage, in_flight=in_flight)
        dt = max_over_ranks(time.perf_counter() - t0)
        assert seen[0] == n_batches * BATCH
        return world * n_batches * BATCH / dt, m.kernel_launches() - l0

    loop_batches = bps * steps
    loop_i16, loop_launches = time_loop(Xi, loop_batches, args.in_flight)
    loop_f32, _ = time_loop(X, loop_batches, args.in_flight)
    loop_pageable, _ = time_loop(X, max(bps, loop_batches // 4), args.in_flight, pageable=True)
    loop_lockstep, lockstep_launches = time_loop(Xi, min(loop_batches, 300), 1)

    # ---- the same end-to-end call with the first-choice variant decision fused behind the heads
    #      (clairb_predict_decide), and the reference-equivalent Python restatement timed on a small sample ----
    ref_bases = (np.arange(sites) % 4).astype(np.uint8)
    dec_steps = max(2, steps // 4)
    for _ in range(2):
        _, dec = m.predict_and_decide_packed(Xi, ref_bases)
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(dec_steps):
        _, dec = m.predict_and_decide_packed(Xi, ref_bases)
    torch.cuda.synchronize()
    e2e_dec_s = max_over_ranks(time.perf_counter() - t0)
    # the batched VCF output stage behind it (clair_b200.output.BatchOutput in place of call_var.batch_output): rows of
    # reference / SNP calls formatted natively, insertion / deletion calls through the indel-base helpers (stand-ins here)
    output_info = None
    if rank == 0:
        import types
        from clair_b200 import output as _output
        nb_out = min(64, bps)
        infos_pool = [[["chr20", str(1000 * b + j + 1), "ACGT"[(b + j) % 4] * 33] for j in range(BATCH)] for b in range(4)]
        helper = types.SimpleNamespace(
            insertion_bases_using=lambda tensor_input, variant_length, contig, position: ("A" * variant_length, variant_length),
            deletion_bases_using=lambda tensor_input, variant_length, contig, position, reference_sequence: (("CGTA" * 5)[:variant_length], variant_length),
            insertion_bases_using_pysam_using=lambda **kw: "C" * kw["minimum_insertion_length"],
            print_debug_message=lambda *a: None)
        out_bytes = [0]
        helper.output = lambda s: out_bytes.__setitem__(0, out_bytes[0] + len(s))
        cfg = types.SimpleNamespace(is_show_reference=True, is_debug=False, is_haploid_precision_mode_enabled=False,
                                    is_haploid_sensitive_mode_enabled=False, is_output_for_ensemble=False, quality_score_for_pass=None)
        stage = _output.BatchOutput(cfg, helper, fallback=lambda *a: None)
        preds = []
        for b in range(nb_out):
            xb = Xi[b * BATCH:(b + 1) * BATCH]
            infos_b = infos_pool[b % 4]
            preds.append((xb, infos_b) + m.predict_and_decide(xb, _decision_codes(infos_b)))
        t0 = time.perf_counter()
        for xb, infos_b, pred, dec_b in preds:
            stage((xb, infos_b), pred, dec_b)
        dt_out = time.perf_counter() - t0
        # a random-init network answers "insertion + deletion" almost everywhere; a call set is reference / SNP calls with a
        # few indels.  Heads shaped like that (peaked ACGT-pair label, both lengths zero, 1 site in 50 an indel) -> records from
        # the device's decision kernel (clairb_decide) -> the same stage
        rng = np.random.default_rng(5)
        z = rng.normal(0, 1.0, (nb_out * BATCH, 90)).astype(np.float32)
        z[np.arange(len(z)), rng.integers(0, 10, len(z))] += 9.0
        z[:, 24 + 16] += 8.0
        z[:, 57 + 16] += 8.0
        indel = rng.random(len(z)) < 0.02
        z[indel, 15] += 14.0
        z[indel, 24 + 18] += 12.0
        z[indel, 57 + 18] += 12.0
        Pc = np.empty_like(z)
        for a, b in ((0, 21), (21, 24), (24, 57), (57, 90)):
            ez = np.exp(z[:, a:b] - z[:, a:b].max(1, keepdims=True))
            Pc[:, a:b] = ez / ez.sum(1, keepdims=True)
        stage_c = _output.BatchOutput(cfg, helper, fallback=lambda *a: None)
        calls = []
        for b in range(nb_out):
            sl = slice(b * BATCH, (b + 1) * BATCH)
            infos_b = infos_pool[b % 4]
            pb = Pc[sl]
            calls.append((Xi[sl], infos_b, [pb[:, 0:21], pb[:, 21:24], pb[:, 24:57], pb[:, 57:90]],
                          m.decide(pb, _decision_codes(infos_b), Xi[sl])))
        t0 = time.perf_counter()
        for xb, infos_b, pred, dec_b in calls:
            stage_c((xb, infos_b), pred, dec_b)
        dt_call = time.perf_counter() - t0
        output_info = {"value": nb_out * BATCH / dt_call, "unit": "sites/s", "sites": nb_out * BATCH,
                       "rows_native": stage_c.fast_rows, "rows_python": stage_c.slow_rows, "sites_to_fallback": stage_c.fallback_sites,
                       "what": "call-like heads (reference / SNP calls, 2 % insertions), every site prints a row (show reference calls)",
                       "random_init_model": {"value": nb_out * BATCH / dt_out, "rows_native": stage.fast_rows, "rows_python": stage.slow_rows,
                                             "note": "the bench's random-init network calls 'insertion + deletion' on 80 % of the sites: the "
                                                     "per-site Python path (indel-base helper calls) carries them"},
                       "note": "BatchOutput on the device's decision records, one host thread + 4 formatter threads; the reference's "
                               "batch_output spends ~1 ms of Python per site (SURVEY.md 8f)"}
    decision_info = None
    if rank == 0:
        decision_info = {"e2e_with_decision": world * sites * dec_steps / e2e_dec_s, "unit": "sites/s",
                         "extra_d2h_bytes_per_step": sites * 32, "extra_h2d_bytes_per_step": sites,
                         "categories_seen": np.bincount(dec.category, minlength=10).tolist(),
                         "note": "forward + decide_sites kernel per chunk (call_var.py:589-690, 732-760), int16 transport"}

    # ---- training step (BASELINE.json configs[4]: batch 512 per GPU, data-parallel): outside the headline's
    #      timed region; every rank takes part (NCCL all-reduce of the gradient buffer) ----
    train_info = None
    if not args.no_train:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import train_bench
        try:
            train_info = train_bench.device_part(512, steps=12, warm=3, device=local_rank, data_parallel=world > 1)
        except Exception as exc:                       # a side stage must not take the headline line down with it
            train_info = {"error": "%s: %s" % (type(exc).__name__, exc)}
    if rank == 0:
        peaks = measured_peaks()
        prof = {p["kernel"]: p for p in profile}
        dom = max(profile, key=lambda p: p["ms"]) if profile else None
        roofline = None
        if dom is not None:
            sites_per_launch = sites * steps / dom["launches"]
            flop = KERNEL_FLOP_PER_SITE.get(dom["kernel"], 0) * sites_per_launch
            avg_ms = dom["ms"] / dom["launches"]
            achieved = flop / (avg_ms * 1e-3) / 1e12
            traffic, traffic_file = measured_traffic(dom["kernel"], sites_per_launch)
            roofline = {"bound": "tensor", "kernel": dom["kernel"], "achieved": achieved,
                        "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_sustained"],
                        "traffic": traffic,
                        "traffic_source": "ncu dram__bytes_read+write per launch (profiles/%s), scaled by sites" % traffic_file,
                        "algorithmic_bytes": KERNEL_BYTES_PER_SITE.get(dom["kernel"], 0) * sites_per_launch,
                        "peak_source": peaks["source"] + " (sustained bf16; timed region %.1f s)" % (ms * 1e-3),
                        "frac_of_burst_peak": achieved / peaks["bf16_burst"],
                        "avg_launch_ms": avg_ms, "sites_per_launch": sites_per_launch,
                        "whole_path_frac": value / world * FLOP_PER_SITE / 1e12 / peaks["bf16_sustained"],
                        "kernel_share": {k: v["ms"] / max(1e-9, sum(p["ms"] for p in profile)) for k, v in prof.items()},
                        "kernel_frac": {k: (KERNEL_FLOP_PER_SITE.get(k, 0) * sites * steps / max(1e-9, v["ms"] * 1e-3) / 1e12) / peaks["bf16_sustained"]
                                        for k, v in prof.items()},
                        "note": "logit tolerance 1e-4 needs the 3-term fp16 split (3 MMAs per algorithmic MAC): attainable "
                                "ceiling is 1/3 of peak, i.e. frac 0.333 = tensor pipe saturated"}
        # ---- CreateTensor stage: alignments -> tensors on the device, outside the timed region ----
        ct_info = ct_ctx = None
        if world == 1 and not args.no_create_tensor:
            sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools"))
            import ct_bench
            launches_before_ct = m.kernel_launches()
            try:
                ct_info, ct_ctx = ct_bench.device_part(m, 2000000, 5, True, measured_peaks()["hbm_gbs"])
                ct_info["gpu_launches"] = m.kernel_launches() - launches_before_ct
            except Exception as exc:      # a side stage must not take the headline line down with it
                ct_info, ct_ctx = {"error": "%s: %s" % (type(exc).__name__, exc)}, None
        cpu = pinning = None
        if world == 1 and not args.no_cpu_baseline:
            # the only place this arm touches oracle/: the CPU port is timed on a bounded sample AND serves as the
            # checker of what was just measured (same sites, same weights)
            # every worker owns 25 distinct synthetic predict-batches (seeds below): 100 k distinct sites on a 16-core box,
            # the "100 k-site sample" of SURVEY.md 8d's parity gate
            n_own = 25
            seeds = [20240607 + 100 + i for i in range(n_own)]
            arm = CpuArm(weights, seeds, check_batch=np.array(X[:BATCH]))
            try:
                dt1 = arm.run(1)
                per_worker = args.cpu_baseline_batches or max(2, min(n_own, int(12.0 / max(dt1, 1e-3))))    # about 12 s of CPU work
                dt, kept = arm.run(per_worker, keep=True)
            finally:
                arm.close()
            # the same sites through the device: worker r ran pool[0] in the calibration run, then pool[(1 + j) % n_own] for j < per_worker
            sample_sites = worst = flips = flips_outside_margin = 0
            for r in range(arm.procs):
                for j in range(per_worker):
                    idx = (1 + j) % n_own
                    if r == 0 and idx == 0:
                        continue                       # worker 0's first batch is the check batch, compared above
                    xb = synth.synthetic_tensors(BATCH, seed=seeds[idx] + 1000 * r)
                    got = m.predict_packed(xb)
                    want = kept[r][j]
                    worst = max(worst, float(np.abs(got - want).max()))
                    for a, b in ((0, 21), (21, 24), (24, 57), (57, 90)):
                        diff = np.flatnonzero(got[:, a:b].argmax(1) != want[:, a:b].argmax(1))
                        flips += int(diff.size)
                        top2 = np.sort(want[diff, a:b], axis=1)[:, -2:]
                        flips_outside_margin += int(((top2[:, 1] - top2[:, 0]) > 2e-4).sum())
                    sample_sites += BATCH
            assert worst <= 1e-4, "GPU result differs from the CPU port on the parity sample: %g" % worst
            assert flips_outside_margin == 0, "arg-max differs from the CPU port where its two best classes are more than 2e-4 apart"
            parity_sample = {"sites": sample_sites, "max_abs_prob_diff": worst, "argmax_differences": flips,
                             "argmax_differences_outside_2e-4_margin": flips_outside_margin,
                             "note": "distinct synthetic sites computed by the CPU port (fp32) and by the device; an arg-max may "
                                     "only differ where the port's own two best classes are within 2e-4 of each other"}
            v = arm.procs * per_worker * BATCH / dt
            model, cores = cpu_info()
            cpu_first = arm.first                  # worker 0's answer for the first BATCH sites of this run's pool
            err = float(np.abs(out[:BATCH] - cpu_first).max())
            assert err <= 1e-4, "GPU result differs from the CPU port of the reference: %g" % err
            for a, b in ((0, 21), (21, 24), (24, 57), (57, 90)):
                assert (out[:BATCH, a:b].argmax(1) == cpu_first[:, a:b].argmax(1)).all(), "arg-max mismatch vs the CPU port"
            from oracle import decision_oracle as DO
            ns = 300
            t0 = time.perf_counter()
            want, want_p, _ = DO.decide(out[:ns], ref_bases[:ns])
            cpu_dec = ns / (time.perf_counter() - t0)
            got = np.stack([dec.category, dec.len1, dec.len2, dec.aux], axis=1)[:ns]
            assert np.array_equal(got, want) and np.array_equal(dec.max_probability[:ns], want_p), "decision parity failed"
            for i in range(0, ns, 7):                  # record words 6 / 7 (quality score) against the restatement
                assert dec.quality[i] == DO.quality_of_first_choice(out[i], int(want[i, 0]), int(want[i, 3])), "quality parity failed"
            decision_info["cpu_python_restatement_sites_per_s"] = cpu_dec
            decision_info["cpu_sample"] = "%d sites, 1 core; device records bit-exact on them" % ns
            if ct_ctx is not None:
                try:
                    ct_bench.cpu_part(ct_info, ct_ctx)
                except Exception as exc:
                    ct_info["cpu_oracle"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
            cpu = {"value": v, "unit": "sites/s", "cores": arm.procs * arm.threads, "kind": "port",
                   "sample": "%d processes x %d threads x %d predict-batches x %d sites of the same workload, %.1f s" % (
                       arm.procs, arm.threads, per_worker, BATCH, dt),
                   "parity_vs_gpu": {"max_abs_prob_diff": err, "argmax_identical": True, "sites": BATCH},
                   "parity_sample": parity_sample,
                   "cpu_model": model, "host_cores": cores}
            pinning = oracle_pinning(m, weights, np.array(X[:256]))
            if train_info is not None and "error" not in train_info:
                try:
                    train_bench.cpu_part(train_info)
                except Exception as exc:
                    train_info["cpu_oracle"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
        engine = os.environ.get("CLAIRB_ENGINE", "default")
        print(json.dumps({
            "metric": "candidate-sites/sec", "value": value, "unit": "sites/s", "n_gpus": world,
            "steps": steps, "warmup": warm, "ms_per_step": ms / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (fp16 hi/lo split operands, f32 accumulate)", "data": "synthetic",
            "config": {"workload": WORKLOAD % world, "batch": BATCH, "batches_per_step": bps, "sites_per_step": sites,
                       "distinct_batches": distinct,
                       "l2": "inputs larger than L2: %.0f MB of fp32 input re-read per step" % (sites * 4224 / 1e6),
                       "weights": "random-init ONT-shape, seed 1234", "engine": engine,
                       "parallelism": "sites sharded over %d GPU(s), one NCCL gather of [sites,90] per step" % world},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "sites/s", "h2d_bytes_per_step": sites * 2112,
                    "d2h_bytes_per_step": sites * 360 * (world if world > 1 else 1), "gpu_launches": e2e_launches,
                    "note": ("one Clair.predict call per step over the pool" if world == 1 else
                             "per rank: host input -> device rows (clairb_predict_async_to_device, %d slices queued per step), NCCL gather of "
                             "every slice to rank 0 behind the forward of the next ones, rank 0 copies N x [sites,90] to host" % E2E_SLICES)
                            + "; pinned host input in the int16 transport (same integer counts as float32, bit-identical output)"},
            "e2e_f32": {"value": e2e_f32_value, "unit": "sites/s", "h2d_bytes_per_step": sites * 4224,
                        "d2h_bytes_per_step": sites * 360, "h2d_gbs_measured": h2d_gbs,
                        "h2d_ceiling_sites_per_s": world * h2d_gbs * 1e9 / 4224,
                        "note": "float32 input (the reference generator's dtype): bounded by the pinned host->device copy"},
            "loop_e2e": {"value": loop_i16, "unit": "sites/s", "batch": BATCH, "in_flight": args.in_flight, "batches": loop_batches,
                         "gpu_launches": loop_launches, "frac_of_e2e": loop_i16 / e2e_value,
                         "f32_pinned": loop_f32, "f32_pageable": loop_pageable,
                         "note": "generator -> one predict per 1000-site batch (run_batches over predict_async) -> no-op output "
                                 "stage; headline = int16 pinned batches"},
            "loop_lockstep": {"value": loop_lockstep, "unit": "sites/s", "gpu_launches": lockstep_launches,
                              "note": "the same loop with exactly one predict in flight (the reference's own loop structure, "
                                      "clair/call_var.py:1327-1352): 1000 sites = 8 of 74 CTA pairs per call"},
            "multi_gpu_check": multi_gpu_check,
            "decision_stage": decision_info,
            "output_stage": output_info,
            "create_tensor_stage": ct_info,
            "train_stage": train_info,
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "oracle_pinning": pinning,
            "kernels": profile,
        }))
    m.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
