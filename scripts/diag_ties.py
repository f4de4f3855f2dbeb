"""Diagnostic (GPU): where mdn_instance_mask_resize, torchvision on the CPU and torchvision on CUDA disagree, and the float64 value there."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import torch.nn.functional as F
import common
from mdn_sfm_b200 import loss_utils
from oracle import restate

for B, seed in ((4, 42), (4, 13), (12, 53)):
    opt, batch = common.make(B, 192, 640, seed=seed)
    inst = batch[4]
    size = (192, 640)
    cpu = restate.resized_instance_mask(inst, size)[:, 0]
    dev_inst = [{"instances": d["instances"].to("cuda")} for d in inst]
    cuda = restate.resized_instance_mask(dev_inst, size)[:, 0].cpu()
    ours = loss_utils.instance_masks_u8(dev_inst, [size], "cuda")[0].cpu().long()
    full = restate.get_batch_instance_mask(inst)[:, :1]
    v64 = F.interpolate(full.double(), size=size, mode="bilinear", align_corners=False, antialias=True)[:, 0]
    v32c = F.interpolate(full.float(), size=size, mode="bilinear", align_corners=False, antialias=True)[:, 0]
    v32g = F.interpolate(full.float().cuda(), size=size, mode="bilinear", align_corners=False, antialias=True)[:, 0].cpu()
    print("B", B, "seed", seed, "cpu!=cuda", int((cpu != cuda).sum()), "ours!=cpu", int((ours != cpu).sum()), "ours!=cuda", int((ours != cuda).sum()))
    bad = ours != cuda
    idx = bad.nonzero()
    for i in idx[:12]:
        b, y, x = [int(v) for v in i]
        print("   px", b, y, x, "v64 %.10f v32cpu %.10f v32cuda %.10f" % (v64[b, y, x], v32c[b, y, x], v32g[b, y, x]), "cpu", int(cpu[b, y, x]), "cuda", int(cuda[b, y, x]), "ours", int(ours[b, y, x]))
