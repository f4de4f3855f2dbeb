"""GPU timeline of one graph replay of the bench step (kernels, memcpys, memsets with start offsets and gaps), from
torch.profiler / CUPTI.  Experiment tooling: python scripts/timeline.py [--bwd 0|1] [--steps-per-graph N] [--mode T|DC|...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from mdn_sfm_b200 import synthetic
from mdn_sfm_b200.loss_functions import Loss


def arg(name, default):
    return sys.argv[sys.argv.index(name) + 1] if name in sys.argv else default


bwd = int(arg("--bwd", "1"))
n_per = int(arg("--steps-per-graph", "4"))
mode = arg("--mode", "T")
with_inst = mode in ("DS", "DC")
B, H, W, scales = 12, 192, 640, (0, 1, 2, 3)
opt = synthetic.default_opt(B, H, W)
sets = []
for i in range(4):
    inputs, flows, mobiles, cams, inst = synthetic.make_batch(B, H, W, scales=scales, seed=42 + i, flow_std=0.05, device="cuda", with_instances=with_inst)
    g = lambda d: {k: v.requires_grad_(True) for k, v in d.items()}
    sets.append((inputs, g(flows), g(mobiles), g(cams), inst))
loss = Loss(opt, no_ssim=False, mode=mode, photometric=True)


def step(i):
    inputs, flows, mobiles, cams, inst = sets[i % 4]
    for d in (flows, mobiles, cams):
        for v in d.values():
            v.grad = None
    _, losses = loss(inputs, [-1, 1], flows, mobiles, inst, list(scales), cams)
    if bwd:
        losses["loss"].backward()


for i in range(8):
    step(i)
torch.cuda.synchronize()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for i in range(4):
        step(i)
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for i in range(n_per):
        step(i)
for _ in range(20):
    g.replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
prev_end = None
for e in evs:
    gap = (e.time_range.start - prev_end) if prev_end is not None else 0.0
    print("%9.1f us  +%6.1f gap  %7.1f us  [stream %s] %s" % (e.time_range.start - t0, gap, e.time_range.end - e.time_range.start, getattr(e, "device_resource_id", "?"), e.name[:60]))
    prev_end = e.time_range.end
print("total %.1f us for %d steps" % (evs[-1].time_range.end - t0, 3 * n_per))
