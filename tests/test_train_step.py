"""the DDP train-step harness (host logic) on CPU -- gloo, world size 2, the oracle as the loss.

One optimisation step of the mobile decoder on two ranks (batch sharded, DDP gradient all-reduce, clip, Adam) must
leave every rank with the parameters a single process gets from the global batch: the loss path needs no collective
of its own, and the harness adds none beyond DDP's and the logging all-reduce."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GB, H, W = 4, 32, 64


def _build(opt, seed=3):
    from mdn_sfm_b200.train_step import StandInNets, TrainStep
    from oracle import restate
    torch.manual_seed(seed)
    nets = StandInNets(width=8)

    def oracle_loss(inputs, ids, flows, mobiles, inst, scales, cams):
        return restate.loss_forward(opt, inputs, ids, flows, mobiles, inst, scales, cams, mode="T", photometric=True, ssim_on=True)

    return TrainStep(opt, nets=nets, loss_module=oracle_loss, device="cpu", lr=1e-3, clip_grad=1.0)


def _inputs(lo, hi):
    from mdn_sfm_b200 import synthetic
    inputs, _, _, _, _ = synthetic.make_batch(GB, H, W, seed=23, with_instances=False)
    return {k: v[lo:hi].contiguous() for k, v in inputs.items()}


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    from mdn_sfm_b200 import distributed as D
    from mdn_sfm_b200 import synthetic
    assert D.init("gloo") == world
    lo, hi = D.shard_batch(GB, rank, world)
    ts = _build(synthetic.default_opt(hi - lo, H, W))
    losses = ts.step(_inputs(lo, hi))
    logged = ts.log_losses(losses)
    if rank == 0:
        q.put(([p.detach().numpy().copy() for p in ts.nets.mobile_decoder.parameters()], logged))   # by value
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_step_equals_global_batch_step():
    sys.path.insert(0, ROOT)
    from mdn_sfm_b200 import synthetic
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 23400 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    params2, logged = q.get(timeout=600)
    params2 = [torch.from_numpy(a) for a in params2]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ts = _build(synthetic.default_opt(GB, H, W))
    before = [p.detach().clone() for p in ts.nets.mobile_decoder.parameters()]
    losses = ts.step(_inputs(0, GB))
    after = list(ts.nets.mobile_decoder.parameters())
    assert any(not torch.equal(a, b) for a, b in zip(after, before)), "the step must move the trainable net"
    assert all(not p.requires_grad for p in ts.nets.flownet.parameters()), "flow / pose nets stay frozen (trainer.py:181-186)"
    # Adam's first step moves every weight by ~lr * sign(g): compare the UPDATES, relative to their size
    for a, b, c in zip(after, params2, before):
        assert float((a - b).abs().max()) <= 1e-3 * float((a - c).abs().max()) + 1e-9
    assert logged["loss"] == pytest.approx(float(losses["loss"].detach()), rel=1e-5)


def test_standin_nets_have_the_reference_interfaces():
    from mdn_sfm_b200.train_step import StandInNets
    nets = StandInNets(width=8)
    tgt, ref = torch.randn(2, 3, 64, 96), torch.randn(2, 3, 64, 96)
    flow, feats = nets.flownet(tgt, ref, frame_id=-1)
    aa, t = nets.posenet(tgt, ref)
    mob = nets.mobile_decoder(feats, aa, t, frame_id=-1)
    for s in range(4):
        assert tuple(flow[("flow", -1, s)].shape) == (2, 2, 64 >> s, 96 >> s)
        assert tuple(mob[("mobile", -1, s)].shape) == (2, 1, 64 >> s, 96 >> s)
        assert 0 < float(mob[("mobile", -1, s)].min()) and float(mob[("mobile", -1, s)].max()) < 1
    assert tuple(aa.shape) == (2, 1, 1, 3) and tuple(t.shape) == (2, 1, 1, 3)
