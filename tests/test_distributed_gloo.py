"""world_size-2 gloo test of the batch-sharded loss path (host logic; the arithmetic is the CPU oracle's).

Checks the section-8e claim the multi-GPU design rests on: with equal shards, the mean over ranks of the per-rank
loss equals the loss of the global batch and the rank-averaged gradients equal the global gradients -- so the loss
kernels need no collective -- plus the helper collectives (`mean_losses`, `max_over_ranks`, `shard_batch`).
"""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, mode, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    import common
    from mdn_sfm_b200 import distributed as D
    from mdn_sfm_b200 import synthetic
    assert D.init("gloo") == world
    GB, H, W = 4, 32, 64
    opt_g, batch = common.make(GB, H, W, seed=17)
    lo, hi = D.shard_batch(GB, rank, world)
    inputs, flows, mobiles, cams, inst = batch
    sl = lambda d: {k: v[lo:hi].contiguous() for k, v in d.items()}
    local = (sl(inputs), sl(flows), sl(mobiles), sl(cams), inst[lo:hi])
    opt_l = synthetic.default_opt(hi - lo, H, W)
    _, losses, f, m, _ = common.oracle_run(opt_l, local, mode, True, True)
    mean = D.mean_losses(losses)
    # DDP-style gradient averaging of a per-sample tensor == slice of the global gradient / 1 (already a global mean)
    g = f[("flow", 1, 0)].grad / world
    gathered = [torch.zeros_like(g) for _ in range(world)]
    dist.all_gather(gathered, g)
    slow = D.max_over_ranks(1.0 + rank)
    if rank == 0:
        _, gl, gf, _, _ = common.oracle_run(opt_g, batch, mode, True, True)
        # numpy arrays travel BY VALUE: a torch tensor on a Queue is a shared-memory handle that the parent can only open
        # while this process is alive, and the worker may exit first (seen as FileNotFoundError in the parent's q.get)
        q.put(({k: float(v) for k, v in mean.items()}, {k: float(gl[k].detach()) for k in mean}, torch.cat(gathered, 0).numpy().copy(),
               gf[("flow", 1, 0)].grad.numpy().copy(), float(slow)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["T", "SN"])
def test_rank_mean_equals_global_batch(mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    mean, glob, g_sharded, g_global, slow = q.get(timeout=300)
    g_sharded, g_global = torch.from_numpy(g_sharded), torch.from_numpy(g_global)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for k in mean:
        assert mean[k] == pytest.approx(glob[k], rel=1e-5), k
    assert ((g_sharded - g_global).abs().max() / g_global.abs().max()).item() < 1e-5
    assert slow == 2.0


def test_shard_helpers():
    from mdn_sfm_b200 import distributed as D
    assert D.shard_batch(24, 1, 2) == (12, 24)
    assert D.shard_seed(42, 3) == 45
    with pytest.raises(ValueError):
        D.shard_batch(10, 0, 4)
    assert D.max_over_ranks(3.5) == 3.5
