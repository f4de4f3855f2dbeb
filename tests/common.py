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


class tie_ruling:
    """DS / DC only.  The reference resizes the integer instance mask with torchvision's antialiased bilinear `Resize` (fp32)
    and rounds.  Where the resized value is a 0.5 tie -- exactly (rectangle edges of the synthetic Detectron2 masks land on
    such ties: ~17 pixels per 375x1242 -> 192x640 sample) or to within the fp32 error of ATen's own filter (measured on
    B200: its CPU and CUDA kernels return 0.5000051 and 0.4999982 for a float64 value of 0.4999940) -- the reference's own
    CPU and CUDA kernels round to different sides.  `mdn_instance_mask_resize` replays the CPU kernel bit for bit
    (scripts/diag_ties.py: `ours != cpu` is 0), so against the oracle run on CUDA a few pixels per batch differ, and a
    flipped mask pixel changes that pixel's gradients by O(1).

    Inside this context `restate.resized_instance_mask` returns the reference's own mask with the PRODUCT's value
    substituted at exactly those pixels where the two disagree -- after proving for each of them that it is such a tie:
    |float64 value - 0.5| < 1e-6, or < 1e-4 AND the product equals the reference's own CPU kernel there.  Anything else
    fails the test.  Comparisons then run with no exempted pixel at all.  `.substituted` counts the pixels."""

    def __init__(self, product_device, library=None):
        self.dev, self.library, self.substituted = product_device, library, 0

    def __enter__(self):
        import torch.nn.functional as F
        from mdn_sfm_b200 import loss_utils
        self._orig = orig = restate.resized_instance_mask
        cache = {}

        def ruled(instances_info, size):
            size = tuple(int(v) for v in size)
            key = (id(instances_info), size)
            if key in cache:
                return cache[key]
            ref = orig(instances_info, size)                                   # (B or 1, 3, h, w) int64, the oracle's device
            masks = instances_info if isinstance(instances_info, list) else [{"instances": instances_info}]
            moved = [{"instances": d["instances"].to(self.dev)} for d in masks]
            got = loss_utils.instance_masks_u8(moved, [size], self.dev, self.library)[0].to(ref.device)   # (B, h, w) uint8
            bad = got.unsqueeze(1) != ref[:, :1]
            if bool(bad.any()):
                full = restate.get_batch_instance_mask([{"instances": d["instances"].to("cpu")} for d in masks])[:, :1].double()
                v64 = F.interpolate(full, size=size, mode="bilinear", align_corners=False, antialias=True)
                dist = (v64 - 0.5).abs()[:, 0]
                cpu_ref = orig([{"instances": d["instances"].to("cpu")} for d in masks], size)[:, 0]
                proven = (dist < 1e-6) | ((dist < 1e-4) & (got.cpu().long() == cpu_ref))
                assert bool(proven[bad[:, 0].cpu()].all()), ("instance masks differ away from a 0.5 tie", size, int(bad.sum()),
                                                             float(dist[bad[:, 0].cpu()].max()))
                self.substituted += int(bad.sum())
                ref = torch.where(bad.expand_as(ref), got.unsqueeze(1).expand_as(ref).to(ref.dtype), ref)
            cache[key] = ref
            return ref

        restate.resized_instance_mask = ruled
        return self

    def __exit__(self, *exc):
        restate.resized_instance_mask = self._orig
        return False


def oracle_run(opt, batch, mode, photo, ssim_on, device="cpu", pose_grad=False, product_device=None, library=None,
               rule_ties=True, padding_mode="zeros"):
    """product_device / library: where (and with which build) the product's instance masks are made for the DS / DC tie
    ruling (see tie_ruling); default = the oracle's own device.  rule_ties=False: the oracle exactly as it is (comparisons
    against the reference's golden fixtures, where no product is involved)."""
    if rule_ties and mode in ("DS", "DC") and batch[4] is not None:
        with tie_ruling(product_device or device, library):
            return _oracle_run(opt, batch, mode, photo, ssim_on, device, pose_grad, padding_mode)
    return _oracle_run(opt, batch, mode, photo, ssim_on, device, pose_grad, padding_mode)


def _oracle_run(opt, batch, mode, photo, ssim_on, device="cpu", pose_grad=False, padding_mode="zeros"):
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
                                       photometric=photo, ssim_on=ssim_on, padding_mode=padding_mode)
    losses["loss"].backward()
    return out, losses, f, m, cams


def product_run(opt, batch, mode, photo, ssim_on, device, pose_grad=False, library=None, arith=None, pose_in=True,
                padding_mode="zeros"):
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
    loss = Loss(opt, no_ssim=not ssim_on, padding_mode=padding_mode, mode=mode, photometric=photo, library=library, arith=arith)
    loss.pose_in = pose_in
    out, losses = loss(inputs, [-1, 1], f, m, inst, scales, cams)
    losses["loss"].backward()
    return out, losses, f, m, cams


def compare(ref, got, photo, check_maps=True, fwd_tol=FWD_TOL, grad_tol=GRAD_TOL):
    """No pixel is exempted: DS / DC oracle runs go through tie_ruling."""
    o_r, l_r, f_r, m_r, c_r = ref
    o_g, l_g, f_g, m_g, c_g = got
    for k in ("loss", "epip", "smooth", "consis") + (("photo",) if photo else ()):
        a, b = float(l_r[k]), float(l_g[k])
        assert abs(a - b) <= fwd_tol * max(abs(a), 1e-12), (k, a, b)
    for k in f_r:
        assert rel_max(f_r[k].grad, f_g[k].grad) <= grad_tol, ("d/dflow", k, rel_max(f_r[k].grad, f_g[k].grad))
    for k in m_r:
        assert rel_max(m_r[k].grad, m_g[k].grad) <= grad_tol, ("d/dmobile", k, rel_max(m_r[k].grad, m_g[k].grad))
    for k in c_r:
        if c_r[k].grad is not None:
            assert rel_max(c_r[k].grad, c_g[k].grad) <= grad_tol, ("d/dpose", k, rel_max(c_r[k].grad, c_g[k].grad))
    if check_maps:
        names = ["epipolars", "epipolar_ori", "flows"] + (["warps", "diffs"] if photo else [])
        for name in names:
            assert set(o_r[name].keys()) == set(o_g[name].keys()), name
            for key in o_r[name]:
                assert o_r[name][key].shape == o_g[name][key].shape, (name, key)
                assert rel_max(o_r[name][key], o_g[name][key]) <= fwd_tol, (name, key, rel_max(o_r[name][key], o_g[name][key]))
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
