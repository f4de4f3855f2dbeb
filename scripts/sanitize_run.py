"""Small end-to-end pass over every kernel of the library (for compute-sanitizer memcheck / racecheck / initcheck):
python scripts/sanitize_run.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mdn_sfm_b200 import loss_utils, pyramid, synthetic
from mdn_sfm_b200.loss_functions import Loss

for (B, H, W, scales, mode) in [(2, 64, 128, (0, 1, 2, 3), "T"), (1, 37, 75, (0,), "SN"), (2, 48, 160, (0, 1), "DC"), (1, 23, 45, (0,), "TG")]:
    opt = synthetic.default_opt(B, H, W, threshold=0.8625 if mode == "TG" else 9.22)
    inputs, flows, mobiles, cams, inst = synthetic.make_batch(B, H, W, scales=scales, seed=1, flow_std=0.08, device="cuda",
                                                              with_instances=mode in ("DS", "DC"))
    g = lambda d: {k: v.requires_grad_(True) for k, v in d.items()}
    flows, mobiles, cams = g(flows), g(mobiles), g(cams)
    if inst is not None:
        inst = [{"instances": d["instances"].to("cuda")} for d in inst]
    loss = Loss(opt, no_ssim=False, mode=mode, photometric=True)
    out, losses = loss(inputs, [-1, 1], flows, mobiles, inst, list(scales), cams)
    (losses["loss"] * 0.5).backward()      # non-unit upstream gradient: scale_grads runs for real
    _ = out["epipolars"][(-1, 0)].sum().item(), out["warps"][(1, 0)].sum().item()    # the maps-only launch
    torch.cuda.synchronize()
    print(mode, float(losses["loss"]))
img = torch.rand(2, 3, 48, 80, device="cuda")
print([tuple(t.shape) for t in pyramid.image_pyramid(img, [(24, 40), (12, 20)])])
torch.cuda.synchronize()
print("done")
