"""SSIM map of the standalone kernel vs the oracle in fp32, both against the oracle in float64 (fuzz_gpu_ops cases that exceeded 1e-5)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, common
from mdn_sfm_b200 import layers
from oracle import restate
for (B, h, w, seed) in [(2, 88, 196, 704972), (3, 106, 13, 902161), (2, 19, 298, 400943), (3, 95, 59, 577435)]:
    g = torch.Generator().manual_seed(seed)
    ref = torch.rand(B, 3, h, w, generator=g).cuda()
    x = torch.rand(B, 3, h, w, generator=g).cuda()
    s32 = restate.ssim(x, ref)
    s64 = restate.ssim(x.double(), ref.double())
    sg = layers.SSIM()(x, ref)
    sc = float(s64.abs().max())
    print("%dx%dx%d: product vs oracle fp32 %.3g | vs float64: oracle fp32 %.3g, product %.3g" % (
        B, h, w, common.rel_max(s32, sg), float((s32.double() - s64).abs().max()) / sc, float((sg.double() - s64).abs().max()) / sc))
