"""GPU fuzz of the standalone operators and the data-side entry points on random shapes against the oracle / torchvision:
inverse_warp (all padding modes), SSIM, smooth_loss, binary_image, FlowWarp, get_epipolar_new on point sets, image pyramids
(planar + packed), uint8 frames.  python scripts/fuzz_gpu_ops.py [N] [seed0]"""
import os
import random
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import common
from mdn_sfm_b200 import layers, loss_utils, pyramid, utils
from oracle import restate

DEV = "cuda"


def one(rng, case):
    B, h, w, seed = case["B"], case["h"], case["w"], case["seed"]
    g = torch.Generator().manual_seed(seed)
    ref = torch.rand(B, 3, h, w, generator=g).to(DEV)
    x = torch.rand(B, 3, h, w, generator=g).to(DEV)
    flow = (torch.randn(B, 2, h, w, generator=g) * case["fpx"]).to(DEV)
    pix = restate.create_coords(B, h, w, DEV)
    pad = case["pad"]
    fo = flow.clone().requires_grad_(True)
    wo, vo = restate.inverse_warp(ref, fo, pix, pad)
    (wo * x).sum().backward()
    fg = flow.clone().requires_grad_(True)
    wg, vg = loss_utils.inverse_warp(ref, fg, pix, pad)
    (wg * x).sum().backward()
    assert common.rel_max(wo, wg) < 1e-5, ("warp", common.rel_max(wo, wg))
    assert torch.equal(vo, vg), "valid mask"
    assert common.rel_max(fo.grad, fg.grad) < 1e-4, ("d/dflow", common.rel_max(fo.grad, fg.grad))
    if h >= 3 and w >= 3:
        xo, yo = x.clone().requires_grad_(True), ref.clone().requires_grad_(True)
        wt = flow[:, :1].abs()
        (restate.ssim(xo, yo) * wt).sum().backward()
        xg, yg = x.clone().requires_grad_(True), ref.clone().requires_grad_(True)
        sg = layers.SSIM()(xg, yg)
        (sg * wt).sum().backward()
        assert common.rel_max(restate.ssim(x, ref), sg) < 1e-5, ("ssim", common.rel_max(restate.ssim(x, ref), sg))
        assert common.rel_max(xo.grad, xg.grad) < 1e-4 and common.rel_max(yo.grad, yg.grad) < 1e-4, "ssim grads"
    m = torch.rand(B, 1, h, w, generator=g).to(DEV)
    a, b = float(loss_utils.smooth_loss(x, m)), float(restate.smooth_loss(x, m))
    assert abs(a - b) <= 1e-5 * max(abs(b), 1e-12), ("smooth", a, b)
    thr = rng.choice([0.2, 0.4, 0.5, 0.9])
    assert torch.equal(utils.binary_image(m, thr), restate.binary_image(m, thr)), "binary_image"
    fa, fb = utils.FlowWarp(B, h, w)(flow), restate.flow_warp_grid(flow)
    assert all(torch.equal(p, q) for p, q in zip(fa, fb)), "FlowWarp"
    # point sets (evaluate_flow.py:105-113)
    n = rng.randint(1, 3000)
    p1 = torch.cat([torch.rand(B, 2, n, generator=g) * torch.tensor([w, h]).view(1, 2, 1), torch.ones(B, 1, n)], 1).to(DEV)
    p2 = p1.clone()
    p2[:, :2] += (torch.randn(B, 2, n, generator=g) * 3).to(DEV)
    K = torch.tensor([[0.58 * w, 0, 0.5 * w], [0, 1.92 * h, 0.5 * h], [0, 0, 1]], dtype=torch.float32)
    inv_K = torch.linalg.pinv(K).unsqueeze(0).repeat(B, 1, 1).to(DEV)
    from mdn_sfm_b200 import synthetic
    cam = synthetic.make_pose(torch.randn(B, 1, 1, 3, generator=g) * 0.05, torch.randn(B, 1, 1, 3, generator=g) * 0.2).to(DEV)
    R, t = cam[:, :3, :3], cam[:, :3, -1]
    if B >= 2:      # (B = 1: the reference's gemm / bgemm difference, scripts/diag_batch1.py)
        eo = restate.get_epipolar_new(p1, p2, inv_K, R, t)
        eg = loss_utils.get_epipolar_new(p1, p2, inv_K, R, t)
        assert common.rel_max(eo, eg) < 1e-5, ("points", common.rel_max(eo, eg))
    # pyramids vs torchvision on the CPU (fp32, antialiased bilinear)
    if h >= 8 and w >= 8:
        from torchvision.transforms import Resize
        sizes = [(max(1, h // 2), max(1, w // 2)), (max(1, h // 4), max(1, w // 4))]
        got = pyramid.image_pyramid(ref, sizes)
        pk = pyramid.image_pyramid(ref, sizes, packed=True)
        for s, t_, p_ in zip(sizes, got, pk):
            want = Resize(s)(ref.cpu())
            assert float((t_.cpu() - want).abs().max()) <= 4e-6, ("pyramid", s, float((t_.cpu() - want).abs().max()))
            assert torch.equal(p_[..., :3].permute(0, 3, 1, 2).contiguous(), t_), ("packed pyramid", s)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    fails = 0
    for it in range(n):
        case = dict(B=rng.randint(1, 3), h=rng.randint(2, 150), w=rng.randint(2, 300), seed=rng.randint(0, 10 ** 6),
                    fpx=rng.choice([0.5, 3.0, 9.0, 40.0]), pad=rng.choice(["zeros", "zeros", "border", "reflection"]))
        try:
            one(rng, case)
        except AssertionError as e:
            fails += 1
            print("FAIL", it, case, "->", str(e)[:300], flush=True)
        except Exception as e:
            fails += 1
            print("ERROR", it, case, "->", type(e).__name__, str(e)[:300], flush=True)
            traceback.print_exc(limit=4)
        if (it + 1) % 100 == 0:
            print("...", it + 1, "cases,", fails, "failures", flush=True)
    print("fuzz (ops): %d cases, %d failures" % (n, fails))
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
