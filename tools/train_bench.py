"""Training-step throughput (BASELINE.json configs[4]: batch 512 per GPU, data-parallel): python tools/train_bench.py [batch]
(or under torchrun for the data-parallel step).  bench.py adds the result to its JSON line as `train_stage`; `cpu_part` is the only piece that touches oracle/."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

FLOP_PER_SITE_FORWARD = 40386432            # SURVEY.md 8d; a training step is ~3x that (forward + two backward contractions)


def labels(n, seed):
    rng = np.random.default_rng(seed)
    Y = np.zeros((n, 90), np.float32)
    for lo, k in ((0, 21), (21, 3), (24, 33), (57, 33)):
        Y[np.arange(n), lo + rng.integers(0, k, n)] = 1
    return Y


def device_part(batch=512, steps=8, warm=2, device=0, data_parallel=False):
    import torch
    from clair_b200 import synth, weights as W
    from clair_b200.train import DataParallelTrainer, Trainer
    world = 1
    if data_parallel:
        import torch.distributed as dist
        world = dist.get_world_size()
        t = DataParallelTrainer(device=device, max_batch=batch)
        rank = dist.get_rank()
    else:
        t = Trainer(device=device, max_batch=batch)
        rank = 0
    t.set_weights(W.random_weights(seed=1234))
    # tensors and labels of a step come from page-locked host memory (the staging a loader fills), copied to the device every step
    from clair_b200.model import pinned_empty
    pool = []
    for i in range(4):
        X, Y = pinned_empty((batch, 33, 8, 4), np.int16), pinned_empty((batch, 90), np.float32)
        X[...] = synth.synthetic_tensors(batch, seed=900 + 10 * rank + i).astype(np.int16)
        Y[...] = labels(batch, 10 * rank + i)
        pool.append((X, Y))
    for i in range(warm):
        t.train(*pool[i % 4])
    torch.cuda.synchronize(device)
    if data_parallel:
        dist.barrier()
    l0 = t.kernel_launches()
    t0 = time.perf_counter()
    parts = {"forward_backward": 0.0, "backward_lstm": 0.0, "apply": 0.0}
    first = last = None
    for i in range(steps):
        last = t.train(*pool[i % 4])
        first = last if first is None else first
    torch.cuda.synchronize(device)
    dt = time.perf_counter() - t0
    if data_parallel:
        v = torch.tensor([dt], dtype=torch.float64, device="cuda:%d" % device)
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        dt = float(v.item())
    out = {"workload": "training step, batch %d per GPU x %d GPU(s), synthetic (tensor, label) pairs, int16 tensors and labels from pinned host memory every step" % (batch, world),
           "sites_per_s": world * batch * steps / dt, "ms_per_step": dt / steps * 1e3, "steps": steps,
           "gpu_launches_per_step": (t.kernel_launches() - l0) / steps,
           "tflops_fp32": world * batch * steps * 3 * FLOP_PER_SITE_FORWARD / dt / 1e12,
           "loss_first_step": first, "loss_last_step": last, "grad_norm_last_step": t.grad_norm,
           "gradient_bytes_all_reduced_per_step": t.num_params * 4 if data_parallel else 0,
           "note": "fp32 results throughout: the 33 steps of a direction in ONE launch (clusters of 8 CTAs, h exchanged through distributed "
                   "shared memory), input projections / weight gradients / L4 as 3xTF32 mma.sync GEMMs over all steps, slice-dense "
                   "kernels with a lane per channel"}
    t.close()
    from clair_b200.model import pinned_free
    for X, Y in pool:
        pinned_free(X)
        pinned_free(Y)
    return out


def cpu_part(report, n=32):
    """The float64 autograd restatement on the host (one step on n sites) as the CPU baseline of the row."""
    from clair_b200 import synth, weights as W
    from oracle import train_oracle as TO
    import torch
    X = synth.synthetic_tensors(n, seed=5)
    t0 = time.perf_counter()
    TO.train_step(X, labels(n, 1), W.random_weights(seed=1234), TO.make_masks(n, seed=1))
    dt = time.perf_counter() - t0
    report["cpu_oracle"] = {"sites_per_s": n / dt, "sites": n, "threads": torch.get_num_threads(),
                            "kind": "port (torch autograd, float64; TensorFlow not installable)"}


if __name__ == "__main__":
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    if "RANK" in os.environ:                             # under torchrun: the data-parallel step, one rank per GPU
        import torch
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        r = device_part(batch, steps=12, warm=3, device=local, data_parallel=True)
        if dist.get_rank() == 0:
            print(json.dumps(r))
        dist.destroy_process_group()
    else:
        r = device_part(batch)
        cpu_part(r)
        print(json.dumps(r))
