"""Shared parity machinery: run the product and the oracle on the same seeded inputs and compare.

Tolerances (BASELINE.json north star): forward values 1e-5 relative, gradients 1e-4, binary masks bit exact.
Per-pixel tensors are compared in the max norm relative to the tensor's own max.
"""
import torch

from mdn_sfm_b200 import synthetic
from oracle import restate

FWD_TOL = 1e-5
GRAD_TOL = 1e-4

CASES = [
    # mode, photometric, ssim, disable_min
    ("T", True, True, False),      # BASELINE configs[1]
    ("SN", True, True, False),     # configs[0]
    ("TG", True, True, False),     # configs[2]
    ("DC", False, False, False),   # HEAD
    ("DS", False, False, True),
    ("DC", True, True, True),
    ("TG", True, False, True),
]


def leaf(d):
    return {k: v.clone().requires_grad_(True) for k, v in d.items()}


def rel_max(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / a.abs().max().clamp_min(1e-30)).item()


TIE_PX = 3   # per-pixel maps of the DS / DC modes: pixels exempted from a comparison (see rel_max_but)


def rel_max_but(a, b, k):
    """rel_max over (B,C,h,w) maps ignoring the k worst PIXELS.  DS / DC only: the instance mask the reference resizes
    with torchvision sits on an exact 0.5 tie at about one pixel per million, where its own CPU and CUDA kernels round
    to different sides (tests/test_emu_kernels.py::test_instance_mask_union_and_resize_match_torchvision pins that every
    disagreement IS such a tie); a flipped mask pixel changes the gradients of that one pixel by O(1)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    err = (a - b).abs().amax(dim=1).reshape(-1)
    if k and err.numel() > k:
        err = err.topk(k + 1).values[-1]
    else:
        err = err.max()
    return (err / a.abs().max().clamp_min(1e-30)).item()


def oracle_run(opt, batch, mode, photo, ssim_on, device="cpu", pose_grad=False):
    inputs, flows, mobiles, cams, inst = batch
    mv = lambda d: {k: v.to(device) for k, v in d.items()}
    inputs, cams = mv(inputs), mv(cams)
    if inst is not None:
        inst = [{"instances": d["instances"].to(device)} for d in inst]
    f, m = leaf(mv(flows)), leaf(mv(mobiles))
    if pose_grad:
        cams = leaf(cams)
    H, W = inputs[("color", 0, 0)].shape[-2:]
    scales = sorted({k[2] for k in flows})
    weights = restate.gauss_distance_weight(4, H, W) if mode == "TG" else None
    out, losses = restate.loss_forward(opt, inputs, [-1, 1], f, m, inst, scales, cams, mode=mode, weights=weights,
                                       photometric=photo, ssim_on=ssim_on)
    losses["loss"].backward()
    return out, losses, f, m, cams


def product_run(opt, batch, mode, photo, ssim_on, device, pose_grad=False, library=None, arith=None, pose_in=True):
    arith = arith or ("cuda" if str(device).startswith("cuda") else "cpu")
    from mdn_sfm_b200.loss_functions import Loss
    inputs, flows, mobiles, cams, inst = batch
    mv = lambda d: {k: v.to(device) for k, v in d.items()}
    inputs, cams = mv(inputs), mv(cams)
    if inst is not None:
        inst = [{"instances": d["instances"].to(device)} for d in inst]
    f, m = leaf(mv(flows)), leaf(mv(mobiles))
    if pose_grad:
        cams = leaf(cams)
    scales = sorted({k[2] for k in flows})
    loss = Loss(opt, no_ssim=not ssim_on, mode=mode, photometric=photo, library=library, arith=arith)
    loss.pose_in = pose_in
    out, losses = loss(inputs, [-1, 1], f, m, inst, scales, cams)
    losses["loss"].backward()
    return out, losses, f, m, cams


def compare(ref, got, photo, check_maps=True, fwd_tol=FWD_TOL, grad_tol=GRAD_TOL, tie_px=0):
    """tie_px: pixels per gradient map exempted from the comparison -- TIE_PX for the DS / DC modes (rel_max_but), else 0."""
    o_r, l_r, f_r, m_r, c_r = ref
    o_g, l_g, f_g, m_g, c_g = got
    for k in ("loss", "epip", "smooth", "consis") + (("photo",) if photo else ()):
        a, b = float(l_r[k]), float(l_g[k])
        assert abs(a - b) <= fwd_tol * max(abs(a), 1e-12), (k, a, b)
    for k in f_r:
        assert rel_max_but(f_r[k].grad, f_g[k].grad, tie_px) <= grad_tol, ("d/dflow", k, rel_max(f_r[k].grad, f_g[k].grad))
    for k in m_r:
        assert rel_max_but(m_r[k].grad, m_g[k].grad, tie_px) <= grad_tol, ("d/dmobile", k, rel_max(m_r[k].grad, m_g[k].grad))
    for k in c_r:
        if c_r[k].grad is not None:
            assert rel_max(c_r[k].grad, c_g[k].grad) <= grad_tol, ("d/dpose", k, rel_max(c_r[k].grad, c_g[k].grad))
    if check_maps:
        names = ["epipolars", "epipolar_ori", "flows"] + (["warps", "diffs"] if photo else [])
        for name in names:
            assert set(o_r[name].keys()) == set(o_g[name].keys()), name
            for key in o_r[name]:
                assert o_r[name][key].shape == o_g[name][key].shape, (name, key)
                assert rel_max_but(o_r[name][key], o_g[name][key], tie_px) <= fwd_tol, (name, key, rel_max(o_r[name][key], o_g[name][key]))
        if photo:
            for key in o_r["valids"]:
                assert o_g["valids"][key].dtype == torch.bool
                assert torch.equal(o_r["valids"][key].cpu(), o_g["valids"][key].cpu()), ("valids", key)
        for s in o_r["min_mobiles"]:
            assert torch.equal(o_r["min_mobiles"][s].cpu(), o_g["min_mobiles"][s].cpu())


def make(B, H, W, scales=(0, 1, 2, 3), seed=11, flow_std=0.05, **opt_over):
    opt = synthetic.default_opt(B, H, W, **opt_over)
    batch = synthetic.make_batch(B, H, W, scales=scales, seed=seed, flow_std=flow_std)
    return opt, batch


def assert_identical_runs(a, b, maps=("epipolars", "epipolar_ori")):
    """Two product runs agree BIT FOR BIT: loss scalars, every gradient, the listed per-pixel maps."""
    _, l_a, f_a, m_a, c_a = a
    _, l_b, f_b, m_b, c_b = b
    for k in l_a:
        if torch.is_tensor(l_a[k]):
            assert torch.equal(l_a[k].detach().cpu(), l_b[k].detach().cpu()), k
    for da, db, what in ((f_a, f_b, "d/dflow"), (m_a, m_b, "d/dmobile"), (c_a, c_b, "d/dpose")):
        for k in da:
            if da[k].grad is not None or db[k].grad is not None:
                assert torch.equal(da[k].grad.cpu(), db[k].grad.cpu()), (what, k)
    for name in maps:
        for key in a[0][name]:
            assert torch.equal(a[0][name][key].cpu(), b[0][name][key].cpu()), (name, key)
