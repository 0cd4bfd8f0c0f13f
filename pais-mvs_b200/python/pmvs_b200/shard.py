"""Multi-GPU plumbing of the patch-refinement path: candidate patches are independent, so rank r of W
refines the contiguous shard shard_range(n, r, W); after a pass the ranks all-gather fixed-size converged-patch records
(centre, normal, fitness, drop) so every rank holds the whole round before the serial commit. torch.distributed is only
plumbing here: NCCL on GPUs, gloo in the CPU tests."""
import ctypes as C

import numpy as np

from . import abi

RECORD_DOUBLES = 8      # center 3, normal 3, fitness, drop


def shard_range(n, rank, world):
    """Contiguous balanced partition of n units: the first n % world ranks get one extra."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def pack_records(out_records):
    """PmvsPatchOut array (ctypes or structured numpy) -> float64 [n, RECORD_DOUBLES]."""
    from .scene import PATCH_OUT_DTYPE
    o = np.frombuffer(out_records, dtype=PATCH_OUT_DTYPE) if not isinstance(out_records, np.ndarray) else out_records
    rec = np.empty((len(o), RECORD_DOUBLES), dtype=np.float64)
    rec[:, 0:3] = o["center"]
    rec[:, 3:6] = o["normal"]
    rec[:, 6] = o["fitness"]
    rec[:, 7] = o["drop"]
    return rec


def allgather_records(local, n_total, rank, world):
    """All ranks contribute their shard's records [n_local, 8]; returns the full [n_total, 8] array in patch order.
    Shards may differ in length by one, so they are padded to the longest before the collective."""
    import torch
    import torch.distributed as dist
    per = (n_total + world - 1) // world
    buf = torch.zeros((per, RECORD_DOUBLES), dtype=torch.float64, device=local.device)
    buf[:local.shape[0]] = local
    out = torch.empty((world * per, RECORD_DOUBLES), dtype=torch.float64, device=local.device)
    dist.all_gather_into_tensor(out, buf)
    parts = []
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        parts.append(out[r * per:r * per + (hi - lo)])
    return torch.cat(parts, dim=0)
