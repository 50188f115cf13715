"""Multi-GPU sharding of candidate sites.

Every op of the forward is per-site, so the path shards embarrassingly: weights are replicated,
rank r takes the contiguous slice [r*ceil(n/G), (r+1)*ceil(n/G)) of the sites straight from host
memory (inputs never cross NVLink), and the only exchange is ONE gather of the packed [n_r, 90]
probabilities to rank 0 (the four heads travel as one message and are split into views after).
The reference's own scale-out is independent OS processes per genome chunk
(clair/callVarBamParallel.py:90-119); there is no collective to mirror.
"""
import numpy as np


def shard_bounds(n, world_size, rank):
    """Contiguous slice of rank `rank`; the last ranks may be short or empty."""
    per = -(-n // world_size)
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def gather_packed(local_out, n, rank, world_size, group=None):
    """Gather per-rank [n_r,90] torch tensors to rank 0 -> [n,90] (None on other ranks).

    One collective call: slices are padded to ceil(n/G) rows so a single dist.gather moves them.
    """
    import torch
    import torch.distributed as dist
    per = -(-n // world_size)
    pad = torch.zeros((per, local_out.shape[1]), dtype=local_out.dtype, device=local_out.device)
    pad[: local_out.shape[0]] = local_out
    if world_size == 1:
        return pad[:n]
    bufs = [torch.empty_like(pad) for _ in range(world_size)] if rank == 0 else None
    dist.gather(pad, gather_list=bufs, dst=0, group=group)
    if rank != 0:
        return None
    return torch.cat(bufs, dim=0)[:n]


def sharded_predict(predict_packed, X, rank, world_size, device="cpu", group=None):
    """Run `predict_packed` (ndarray [m,33,8,4] -> ndarray [m,90]) on this rank's slice and gather.

    Returns the full [n,90] ndarray on rank 0 (bit-identical to a single-rank run, since sites
    are independent) and None elsewhere.
    """
    import torch
    n = X.shape[0]
    lo, hi = shard_bounds(n, world_size, rank)
    if hi > lo:
        local = np.asarray(predict_packed(X[lo:hi]), dtype=np.float32)
    else:
        local = np.empty((0, 90), dtype=np.float32)
    t = torch.from_numpy(local).to(device)
    full = gather_packed(t, n, rank, world_size, group)
    return None if full is None else full.cpu().numpy()
