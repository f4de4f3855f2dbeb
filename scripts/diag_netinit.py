"""Prints, for the netinit golden fixtures, the deviation of (a) the CUDA path and (b) the oracle run eagerly on the
same GPU from the fixture written by the reference on the CPU -- i.e. our error next to the reference's own CPU<->GPU noise."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import common
import golden_util as gu
from mdn_sfm_b200 import synthetic


def rel(a, b):
    a = torch.as_tensor(np.asarray(a)).double()
    b = b.detach().double().cpu()
    return ((a - b).abs().max() / a.abs().max().clamp_min(1e-30)).item()


for mode in gu.NETINIT_MODES:
    photo, ssim_on, dmin = gu.NETINIT_MODES[mode]
    (B, H, W), batch = gu.load_netinit_batch()
    opt = synthetic.default_opt(B, H, W, disable_min=dmin)
    z = gu.load_outputs("netinit_" + mode)
    ours = common.product_run(opt, batch, mode, photo, ssim_on, "cuda", pose_grad=True, arith="cpu")
    orc = common.oracle_run(opt, batch, mode, photo, ssim_on, "cuda", pose_grad=True)
    for name, run in (("ours", ours), ("oracle@cuda", orc)):
        out, losses, f, m, c = run
        row = {"loss": abs(float(z["loss_loss"]) - float(losses["loss"].detach())) / abs(float(z["loss_loss"]))}
        for i in (-1, 1):
            row["epip%d" % i] = rel(z["epipolars_" + gu.key((i,))], out["epipolars"][(i, 0)][:, :1])
            row["ori%d" % i] = rel(z["epipolar_ori_" + gu.key((i,))], out["epipolar_ori"][(i, 0)][:, :1])
            row["gcam%d" % i] = rel(z["gcam_" + gu.key((i,))], c[i].grad)
        row["gflow"] = max(rel(z["gflow_" + gu.key(k[1:])], v.grad) for k, v in f.items())
        row["gmob"] = max(rel(z["gmob_" + gu.key(k[1:])], v.grad) for k, v in m.items())
        print("%-3s %-12s " % (mode, name) + " ".join("%s=%.1e" % kv for kv in row.items()))
