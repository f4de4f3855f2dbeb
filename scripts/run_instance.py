"""Runs the instance-mask preparation a few times (for ncu): python scripts/run_instance.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mdn_sfm_b200 import loss_utils, synthetic

g = torch.Generator().manual_seed(5)
inst = [{"instances": d["instances"].to("cuda")} for d in synthetic.make_instances(12, g)]
for _ in range(4):
    out = loss_utils.instance_masks_u8(inst, [(192, 640), (96, 320), (48, 160), (24, 80)], "cuda")
torch.cuda.synchronize()
print("ok", [int(o.sum()) for o in out])
