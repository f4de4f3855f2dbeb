"""Runs / times the instance-mask preparation (for ncu and A/B timing): python scripts/run_instance.py [--time]

--time: CUDA-event time of mdn_instance_mask_union + mdn_instance_mask_resize (B = 12, 375 x 1242 -> the four pyramid
levels) with the fused bit-packed resize (default) and with the two separable passes (MDN_RESIZE_TWO_PASS=1)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mdn_sfm_b200 import loss_utils, synthetic

if "--lib" in sys.argv:   # a tuning build instead of the in-tree library
    from mdn_sfm_b200 import _cabi
    _cabi._lib = _cabi.Library(os.path.abspath(sys.argv[sys.argv.index("--lib") + 1]))

g = torch.Generator().manual_seed(5)
inst = [{"instances": d["instances"].to("cuda")} for d in synthetic.make_instances(12, g)]
sizes = [(192, 640), (96, 320), (48, 160), (24, 80)]
for _ in range(4):
    out = loss_utils.instance_masks_u8(inst, sizes, "cuda")
torch.cuda.synchronize()
print("ok", [int(o.sum()) for o in out])
if "--time" in sys.argv:
    for two_pass in (False, True):
        if two_pass:
            os.environ["MDN_RESIZE_TWO_PASS"] = "1"
        for _ in range(5):
            loss_utils.instance_masks_u8(inst, sizes, "cuda")
        g_ = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_):
            for _ in range(10):
                loss_utils.instance_masks_u8(inst, sizes, "cuda")
        g_.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            g_.replay()
        e1.record()
        torch.cuda.synchronize()
        print("%s: %.1f us per preparation (union + resize to 4 levels)" % ("two separable passes" if two_pass else "fused bit-packed pass", e0.elapsed_time(e1) * 1e3 / 200))
