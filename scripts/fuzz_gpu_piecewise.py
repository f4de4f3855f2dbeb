"""GPU fuzz of the PIECEWISE API (LossModule.epipolar_loss called directly; LossModule.forward + single_mobile_mask_forward +
consistency_loss accumulators) on random ragged shapes against the oracle on the same GPU.  python scripts/fuzz_gpu_piecewise.py [N] [seed0]"""
import os
import random
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import common
from mdn_sfm_b200.loss_functions import LossModule
from oracle import restate

DEV = "cuda"
dev = lambda d: {k: v.to(DEV) for k, v in d.items()}


def close(a, b, tol):
    return abs(float(a) - float(b)) <= tol * max(abs(float(a)), 1e-12)


def epipolar_direct(rng, case):
    B, H, W, mode, seed, fstd = case["B"], case["H"], case["W"], case["mode"], case["seed"], case["fstd"]
    opt, batch = common.make(B, H, W, scales=(0,), seed=seed, flow_std=fstd)
    inputs, flows, mobiles, cams, inst = batch
    inputs, cams = dev(inputs), dev(cams)
    inst = [{"instances": d["instances"].to(DEV)} for d in inst]
    info = inst if case["form"] == "list" else inst[0]["instances"]
    i = case["frame"]
    pix = restate.create_coords(B, H, W, DEV)
    f = (restate.get_scale_factor(B, H, W, DEV) * flows[("flow", i, 0)].to(DEV)).contiguous()
    m = mobiles[("mobile", i, 0)].to(DEV)
    R, t = cams[i][:, :3, :3], cams[i][:, :3, -1]
    weights = restate.gauss_distance_weight(1, H, W)
    fo, mo = f.clone().requires_grad_(True), m.clone().requires_grad_(True)
    with common.tie_ruling(DEV):
        lo, po, eo = restate.epipolar_loss(fo, mo, info, inputs[("inv_K", 0)], R, t, pix, mode=mode, alpha=opt.alpha, w_d2_sim=opt.w_d2_sim,
                                           threshold=opt.threshold, weight=weights[0].to(DEV) if mode == "TG" else None)
    lo.backward()
    fg, mg = f.clone().requires_grad_(True), m.clone().requires_grad_(True)
    lg, pg, eg = LossModule(opt, batch=B, mode=mode).epipolar_loss(fg, mg, info, inputs[("inv_K", 0)], R, t)
    lg.backward()
    ft, gt = (2e-4, 4e-4) if B == 1 else (common.FWD_TOL, common.GRAD_TOL)     # (B = 1: the reference's own floor, scripts/diag_batch1.py)
    assert close(lo, lg, ft), ("loss", float(lo), float(lg))
    assert common.rel_max(po, pg) <= ft, ("post", common.rel_max(po, pg))
    assert common.rel_max(eo, eg) <= ft, ("ori", common.rel_max(eo, eg))
    assert common.rel_max(fo.grad, fg.grad) <= gt, ("d/dflow", common.rel_max(fo.grad, fg.grad))
    assert common.rel_max(mo.grad, mg.grad) <= gt, ("d/dmask", common.rel_max(mo.grad, mg.grad))


def module_forward(rng, case):
    B, H, W, mode, seed, fstd = case["B"], case["H"], case["W"], case["mode"], case["seed"], case["fstd"]
    opt, batch = common.make(B, H, W, scales=(0,), seed=seed, flow_std=fstd)
    inputs, flows, mobiles, cams, inst = batch
    inputs, cams = dev(inputs), dev(cams)
    inst = [{"instances": d["instances"].to(DEV)} for d in inst] if mode in ("DS", "DC") else None
    weights = restate.gauss_distance_weight(1, H, W)

    def run(product):
        fl = {k: v.to(DEV).requires_grad_(True) for k, v in flows.items()}
        mo = {k: v.to(DEV).requires_grad_(True) for k, v in mobiles.items()}
        lm = (LossModule(opt, batch=B, mode=mode, weights=[w.to(DEV) for w in weights]) if product
              else restate.LossModule(opt, mode=mode, weights=[w.to(DEV) for w in weights]))
        lm.consistency_loss(mo[("mobile", -1, 0)], mo[("mobile", 1, 0)], 0)
        shared = mo[("mobile", 1, 0)]
        if product:
            lm(inputs, [-1, 1], fl, shared, inst, cams, 0)
            lm.single_mobile_mask_forward(inputs, -1, fl, mo[("mobile", -1, 0)], inst, cams, 0)
        else:
            for i in (-1, 1):
                lm.frame_terms(inputs, i, fl, shared, inst, cams, 0)
            lm.frame_terms(inputs, -1, fl, mo[("mobile", -1, 0)], inst, cams, 0)
        (lm.losses["epip"] + 0.7 * lm.losses["smooth"] + 0.3 * lm.losses["consis"]).backward()
        return lm, fl, mo

    if inst is not None:
        with common.tie_ruling(DEV):
            olm, fo, mo_ = run(False)
    else:
        olm, fo, mo_ = run(False)
    glm, fg, mg = run(True)
    ft, gt = (2e-4, 4e-4) if B == 1 else (common.FWD_TOL, common.GRAD_TOL)
    for k in ("consis", "epip", "smooth"):
        assert close(olm.losses[k], glm.losses[k], ft), (k, float(olm.losses[k]), float(glm.losses[k]))
    for k in fo:
        assert common.rel_max(fo[k].grad, fg[k].grad) <= gt, ("d/dflow", k, common.rel_max(fo[k].grad, fg[k].grad))
    for k in mo_:
        assert common.rel_max(mo_[k].grad, mg[k].grad) <= gt, ("d/dmobile", k, common.rel_max(mo_[k].grad, mg[k].grad))
    for name in ("epipolars", "epipolar_ori"):
        for key, ref in olm.outputs[name].items():
            assert common.rel_max(ref, glm.outputs[name][key]) <= ft, (name, key, common.rel_max(ref, glm.outputs[name][key]))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    fails = 0
    for it in range(n):
        what = rng.choice(["direct", "module"])
        case = dict(what=what, B=rng.randint(1, 3), H=rng.randint(8, 120), W=rng.randint(8, 260), mode=rng.choice(["SN", "T", "TG", "DS", "DC"]),
                    seed=rng.randint(0, 10 ** 6), fstd=rng.choice([0.01, 0.05, 0.2]), form=rng.choice(["list", "list", "bare"]), frame=rng.choice([-1, 1]))
        if case["form"] == "bare" and what == "direct" and case["mode"] in ("DS", "DC"):
            pass            # a bare Instances broadcasts over the batch
        try:
            (epipolar_direct if what == "direct" else module_forward)(rng, case)
        except AssertionError as e:
            fails += 1
            print("FAIL", it, case, "->", str(e)[:300], flush=True)
        except Exception as e:
            fails += 1
            print("ERROR", it, case, "->", type(e).__name__, str(e)[:300], flush=True)
            traceback.print_exc(limit=4)
        if (it + 1) % 100 == 0:
            print("...", it + 1, "cases,", fails, "failures", flush=True)
    print("fuzz (piecewise): %d cases, %d failures" % (n, fails))
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
