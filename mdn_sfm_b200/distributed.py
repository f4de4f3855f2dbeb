"""Multi-GPU plumbing for the loss path: one process per GPU, batch sharding, no data-path collective.

Every term of the loss is a mean over (B, pixels) of per-sample quantities, so with equal
per-rank batches `mean over ranks of the per-rank loss == loss of the global batch` and DDP's gradient averaging is
the gradient of that global mean.  The loss kernels therefore never communicate; the only collectives are the DDP
gradient all-reduce of the trainable net (outside this package) and the small all-reduce of the loss scalars for
logging, `mean_losses` below.  NCCL over NVLink on GPUs; the same code runs on gloo in the CPU tests.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init(backend=None, device=None):
    """Initialise the default process group from the torchrun environment (RANK / WORLD_SIZE / MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 or dist.is_initialized():
        return world
    backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
    kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
    dist.init_process_group(backend, **kw)
    return world


def shard_seed(base_seed, rank):
    """Per-rank seed of the synthetic input factory (42 + rank)."""
    return base_seed + rank


def shard_batch(global_batch, rank, world):
    """Contiguous [begin, end) slice of a global batch owned by `rank` (equal shards required)."""
    if global_batch % world:
        raise ValueError("global batch %d is not divisible by world size %d" % (global_batch, world))
    per = global_batch // world
    return rank * per, (rank + 1) * per


def mean_losses(losses, keys=("loss", "epip", "smooth", "consis", "photo")):
    """All-reduce(mean) of the detached loss scalars in ONE small collective; returns python floats keyed like `losses`."""
    present = [k for k in keys if k in losses and torch.is_tensor(losses[k])]
    vec = torch.stack([losses[k].detach().float().reshape(()) for k in present])
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM)
        vec = vec / dist.get_world_size()
    return dict(zip(present, vec.tolist()))


def max_over_ranks(value, device="cpu"):
    """Max of a python float over ranks (the timing rule: a multi-GPU step takes as long as its slowest rank)."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
