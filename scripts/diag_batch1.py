"""Is the reference's own CUDA result batch-size dependent?  The oracle's epipolar chain on ONE sample, run at B = 1 and as sample 0
of a batch of two identical samples: bitwise comparison of F, F p1 and the distance (torch.matmul dispatches gemm for one batch and
bgemm for more).  python scripts/diag_batch1.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from mdn_sfm_b200 import synthetic
from oracle import restate

dev = "cuda"
for (H, W, seed, fstd) in [(25, 18, 573554, 0.01), (191, 353, 145788, 0.01), (79, 90, 564232, 0.01), (192, 640, 7, 0.05)]:
    inputs, flows, mobiles, cams, _ = synthetic.make_batch(1, H, W, scales=(0,), seed=seed, flow_std=fstd, with_instances=False)
    inv_K = inputs[("inv_K", 0)].to(dev)
    cam = cams[1].to(dev)
    flow = flows[("flow", 1, 0)].to(dev)
    R, t = cam[:, :3, :3], cam[:, :3, 3]
    pix = restate.create_coords(1, H, W).to(dev) if hasattr(restate, "create_coords") else None
    ys, xs = torch.meshgrid(torch.arange(H, device=dev, dtype=torch.float32), torch.arange(W, device=dev, dtype=torch.float32), indexing="ij")
    p1 = torch.stack([xs, ys, torch.ones_like(xs)], 0).reshape(1, 3, -1)
    fl = flow.reshape(1, 2, -1)
    sc = torch.tensor([W, H], device=dev, dtype=torch.float32).view(1, 2, 1) * 0.5
    p2 = torch.cat([p1[:, :2] + fl * sc, p1[:, 2:]], 1)
    rep = lambda x: x.repeat(2, *([1] * (x.dim() - 1)))
    F1 = restate.fundamental_matrix(inv_K[:, :3, :3], R, t)
    F2 = restate.fundamental_matrix(rep(inv_K[:, :3, :3]), rep(R), rep(t))
    Fp1_1 = torch.matmul(F1, p1)
    Fp1_2 = torch.matmul(F2, rep(p1))
    Fp1_2sameF = torch.matmul(rep(F1), rep(p1))
    d1 = restate.get_epipolar_new(p1, p2, inv_K[:, :3, :3], R, t)
    d2 = restate.get_epipolar_new(rep(p1), rep(p2), rep(inv_K[:, :3, :3]), rep(R), rep(t))
    n1 = (Fp1_1 * p2).sum(1, True)
    n2 = (rep(Fp1_1) * rep(p2)).sum(1, True)
    print("%dx%d: F equal %s | F p1 equal (same F) %s, max rel diff %.3g | sum-of-3 equal %s | distance equal %s, rel max diff %.3g (of max |d| %.3g)" % (
        H, W, torch.equal(F1[0], F2[0]), torch.equal(Fp1_1[0], Fp1_2sameF[0]),
        float(((Fp1_1[0] - Fp1_2sameF[0]).abs() / Fp1_1[0].abs().clamp_min(1e-30)).max()),
        torch.equal(n1[0], n2[0]), torch.equal(d1[0], d2[0]), float((d1[0] - d2[0]).abs().max() / d1[0].abs().max()), float(d1[0].abs().max())))
