"""GPU fuzz of the instance-mask preparation (mdn_instance_mask_union + mdn_instance_mask_resize): random source sizes, 1-4 random
target sizes (down- and up-scaling, factors up to ~50), random densities / instance counts, against torchvision's Resize of the
int64 union on the CPU -- bit for bit except at proven exact 0.5 ties.  python scripts/fuzz_gpu_masks.py [N] [seed0]"""
import os
import random
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn.functional as F

from mdn_sfm_b200 import loss_utils
from mdn_sfm_b200.synthetic import SyntheticInstances
from oracle import restate

DEV = "cuda"


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    fails = ties = 0
    for it in range(n):
        B, H, W = rng.randint(1, 3), rng.randint(4, 400), rng.randint(4, 1300)
        n_inst, dens = rng.randint(1, 4), rng.choice([0.02, 0.1, 0.5, 0.9])
        # (target widths >= 2: for a width-1 output ATen's own CPU kernel returns the same value for every output row -- its vertical
        # pass mis-strides a (.., H, 1) temporary; Resize((25, 1)) != Resize((25, 1))(Resize((25, 112))) there -- nothing to be equal to)
        sizes = [(rng.randint(1, max(1, min(2 * H, 256))), rng.randint(2, max(2, min(2 * W, 700)))) for _ in range(rng.randint(1, 4))]
        case = dict(B=B, H=H, W=W, n_inst=n_inst, dens=dens, sizes=sizes)
        try:
            g = torch.Generator().manual_seed(rng.randint(0, 10 ** 6))
            inst = []
            for _ in range(B):
                masks = torch.rand(n_inst, H, W, generator=g) < dens
                for k in range(n_inst):      # a few filled rectangles: edges are where exact ties live
                    y0, x0 = rng.randint(0, H - 1), rng.randint(0, W - 1)
                    masks[k, y0:y0 + rng.randint(1, H), x0:x0 + rng.randint(1, W)] = rng.random() < 0.7
                inst.append({"instances": SyntheticInstances(masks)})
            got = loss_utils.instance_masks_u8([{"instances": d["instances"].to(DEV)} for d in inst], sizes, DEV)
            full = restate.get_batch_instance_mask(inst)[:, :1].double()
            for g_, s in zip(got, sizes):
                ref = restate.resized_instance_mask(inst, s)[:, 0].to(torch.uint8)
                bad = g_.cpu() != ref
                if bad.any():
                    v64 = F.interpolate(full, size=tuple(s), mode="bilinear", align_corners=False, antialias=True)[:, 0]
                    d = float((v64[bad] - 0.5).abs().max())
                    assert d < 1e-6, ("not a tie", s, int(bad.sum()), d)
                    ties += int(bad.sum())
        except AssertionError as e:
            fails += 1
            print("FAIL", it, case, "->", str(e)[:300], flush=True)
        except Exception as e:
            fails += 1
            print("ERROR", it, case, "->", type(e).__name__, str(e)[:300], flush=True)
            traceback.print_exc(limit=4)
    print("fuzz (masks): %d cases, %d failures, %d pixels at proven exact ties" % (n, fails, ties))
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
