"""Eager-torch restatement of the reference loss path (TEST INFRASTRUCTURE ONLY).

Every function replays, op for op and in the reference's own fp32 order, the
reference function it cites (paths relative to the upstream tree,
``chenluchu/MDN_SfM``).  It is device agnostic: on the build container it runs
on CPU; on the B200 box the same code runs eagerly on ``cuda`` and is the
"reference eager GPU path" comparator.  The third-party arithmetic the reference
relies on (``F.grid_sample``, ``AvgPool2d``, ``ReflectionPad2d``,
``torch.max/min``, torchvision ``Resize``; torch 2.11.0 / torchvision 0.26.0)
is called, not re-derived.

Pinned bit-exactly against the reference imported in place by
``tests/test_oracle_vs_reference.py`` (build container only) and against the
committed fixtures under ``tests/golden``.

The mode switch (SN / T / TG / DS / DC) does not exist upstream as a runtime
flag; each mode is DEFINED here as the composition of reference functions that
SURVEY.md section 8a-M lists (live call sites + commented-out call sites).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F
from torchvision.transforms import Resize

MODES = ("SN", "T", "TG", "DS", "DC")


# --------------------------------------------------------------------------- grids
def create_coords(batch_size, height, width, device="cpu"):
    """loss_utils.py:141-148 / loss_functions.py:150-157 -- (B,2,H,W), ch0 = column, ch1 = row."""
    xs = torch.arange(width, dtype=torch.float32, device=device).view(1, 1, 1, width)
    ys = torch.arange(height, dtype=torch.float32, device=device).view(1, 1, height, 1)
    grid = torch.cat([xs.expand(1, 1, height, width), ys.expand(1, 1, height, width)], 1)
    return grid.repeat(batch_size, 1, 1, 1)


def get_scale_factor(batch_size, height, width, device="cpu"):
    """networks/layers.py:101-103 -- [W,H] broadcast to (B,2,H,W)."""
    sf = torch.tensor([float(width), float(height)], dtype=torch.float32, device=device)
    return sf.view(1, 2, 1, 1).expand(batch_size, 2, height, width)


# --------------------------------------------------------------------------- pose
def rot_from_axisangle(vec):
    """networks/layers.py:59-98 -- Rodrigues, (B,1,3) -> (B,4,4)."""
    angle = torch.norm(vec, 2, 2, True)
    axis = vec / (angle + 1e-7)
    ca, sa = torch.cos(angle), torch.sin(angle)
    C = 1 - ca
    x, y, z = (axis[..., k].unsqueeze(1) for k in range(3))
    xs, ys, zs = x * sa, y * sa, z * sa
    xC, yC, zC = x * C, y * C, z * C
    xyC, yzC, zxC = x * yC, y * zC, z * xC
    rot = torch.zeros((vec.shape[0], 4, 4), device=vec.device)
    rot[:, 0, 0] = torch.squeeze(x * xC + ca)
    rot[:, 0, 1] = torch.squeeze(xyC - zs)
    rot[:, 0, 2] = torch.squeeze(zxC + ys)
    rot[:, 1, 0] = torch.squeeze(xyC + zs)
    rot[:, 1, 1] = torch.squeeze(y * yC + ca)
    rot[:, 1, 2] = torch.squeeze(yzC - xs)
    rot[:, 2, 0] = torch.squeeze(zxC - ys)
    rot[:, 2, 1] = torch.squeeze(yzC + xs)
    rot[:, 2, 2] = torch.squeeze(z * zC + ca)
    rot[:, 3, 3] = 1
    return rot


def get_translation_matrix(tvec):
    """networks/layers.py:43-56."""
    T = torch.zeros(tvec.shape[0], 4, 4, device=tvec.device)
    t = tvec.contiguous().view(-1, 3, 1)
    for k in range(4):
        T[:, k, k] = 1
    T[:, :3, 3, None] = t
    return T


def transformation_from_parameters(axis_angle, translation, invert=False):
    """networks/layers.py:16-40."""
    R = rot_from_axisangle(axis_angle.squeeze(1))
    t = translation.clone().squeeze(1)
    if invert:
        R = R.transpose(1, 2)
        t *= -1
    T = get_translation_matrix(t)
    return torch.matmul(R, T) if invert else torch.matmul(T, R)


# --------------------------------------------------------------------------- epipolar
def fundamental_matrix(inv_K, rotation, translation):
    """loss_utils.py:50-62 -- F = K^-T ((t_x R) K^-1), same association order."""
    t_x = torch.zeros_like(rotation)
    t_x[..., 0, 1] = -translation[..., 2]
    t_x[..., 1, 0] = translation[..., 2]
    t_x[..., 0, 2] = translation[..., 1]
    t_x[..., 2, 0] = -translation[..., 1]
    t_x[..., 1, 2] = -translation[..., 0]
    t_x[..., 2, 1] = translation[..., 0]
    Fm = torch.matmul(t_x, rotation)
    return torch.matmul(torch.transpose(inv_K, -2, -1), torch.matmul(Fm, inv_K))


def get_epipolar_new(p1, p2, inv_K, rotation, translation):
    """loss_utils.py:39-69 -- signed point-to-epipolar-line distance, (B,1,N)."""
    Fm = fundamental_matrix(inv_K, rotation, translation)
    Fp1 = torch.matmul(Fm, p1)
    num = (Fp1 * p2).sum(1, True)
    return num / ((torch.sum(Fp1[:, :2, :] ** 2, dim=1, keepdim=True) + 1e-10).sqrt() + 1e-10)


def post_process_epipolar_1(epipolar_map):
    """loss_utils.py:92-99 -- SN.  Divides its ARGUMENT in place (aliasing quirk kept)."""
    b, c, h, w = epipolar_map.size()
    norms = torch.max(epipolar_map.view(b, -1), dim=1, keepdim=True)[0]
    norms = norms[..., None, None].repeat(1, c, h, w)
    epipolar_map /= norms
    return epipolar_map ** 2


def post_pro_epipolar_weighted(epipolar_map, weight=None, threshold=None):
    """loss_utils.py:81-89 -- T (threshold only) / TG (threshold then weight)."""
    post = epipolar_map.clone()
    if threshold is not None:
        post /= threshold
    if weight is not None:
        post /= weight
    return post ** 2


def get_batch_instance_mask(instances_info):
    """loss_utils.py:102-124 -- int64 {0,1}, (B,3,H,W) for a list, (1,3,H,W) for a bare Instances."""
    if isinstance(instances_info, list):
        m = torch.stack([torch.sum(info["instances"].pred_masks, dim=0, keepdim=True).repeat(3, 1, 1)
                         for info in instances_info], dim=0)
    else:
        m = torch.sum(instances_info.pred_masks, dim=0, keepdim=True).repeat(3, 1, 1).unsqueeze(0)
    mask = torch.zeros_like(m)
    mask[m != 0] = 1
    return mask


def resized_instance_mask(instances_info, size):
    """The `Resize(size)(get_batch_instance_mask(.))` step shared by loss_utils.py:73-75 and :135-137."""
    return Resize(tuple(size))(get_batch_instance_mask(instances_info))


def post_process_epipolar_2(epipolar_map, instances_info):
    """loss_utils.py:127-138 -- DS: instance-mask multiply, result (B,3,h,w)."""
    return resized_instance_mask(instances_info, epipolar_map.size()[2:]) * epipolar_map


def detectron2_similarity_loss(mobile_mask, instances_info):
    """loss_utils.py:72-78 -- DC cross entropy map (B,3,h,w)."""
    mask = resized_instance_mask(instances_info, mobile_mask.size()[2:])
    return -(mask * torch.log(mobile_mask + 1e-10) + (1 - mask) * torch.log(1 - mobile_mask + 1e-10))


def gauss_distance_weight(num_scale, height=128, width=416, sigma1=30, sigma2=120):
    """utils.py:355-379 -- TG weights, float64 arithmetic then fp32, one (1,1,h,w) per scale.

    Vectorised over the reference's double loop with the same expression order
    (rho = 0, so the cross term is an exact 0 and sqrt(1-rho^2) == 1).
    """
    out = []
    for s in range(num_scale):
        num = 2 ** s
        h, w = height // num, width // num
        i = np.arange(h, dtype=np.float64).reshape(h, 1)
        j = np.arange(w, dtype=np.float64).reshape(1, w)
        a = (i - h // 2) ** 2 / (sigma1 / num) ** 2
        b = (j - w // 2) ** 2 / (sigma2 / num) ** 2
        c = 2 * 0 * (i - h // 2) * (j - w // 2) / (sigma1 * sigma2)
        factor = num ** 2 / (2 * np.pi * sigma1 * sigma2 * np.sqrt(1 - 0 ** 2)) / num ** 2
        g = factor * np.exp(-(a + b - c) / (2 * (1 - 0 ** 2)))
        d = 2e5 * (g.max() - g) + 5
        out.append(torch.tensor(d).unsqueeze(0).unsqueeze(0).type(torch.float32))
    return out


def binary_image(x, threshold=0.5):
    """utils.py:100-103."""
    out = torch.zeros_like(x)
    out[x >= threshold] = 1
    return out


def flow_warp_grid(flow):
    """utils.py:289-315 (FlowWarp.forward) -- (pix_coords, normalised grid, valid (B,h,w))."""
    b, _, h, w = flow.shape
    pix = create_coords(b, h, w, flow.device) + flow
    g = pix.permute(0, 2, 3, 1).contiguous()
    g[..., 0] /= w - 1
    g[..., 1] /= h - 1
    g = (g - 0.5) * 2
    return pix, g, g.abs().max(dim=-1)[0] <= 1


# --------------------------------------------------------------------------- warp / photometric
def inverse_warp(ref_img, flow_map, pix_coords, padding_mode="zeros"):
    """loss_utils.py:12-36 -- flow warp + 3-channel bool validity."""
    _, _, h, w = flow_map.size()
    grid = (pix_coords + flow_map).permute(0, 2, 3, 1).contiguous()
    grid[..., 0] /= (w - 1)
    grid[..., 1] /= (h - 1)
    grid = 2 * grid - 1
    warped = F.grid_sample(ref_img, grid, padding_mode=padding_mode, align_corners=True)
    valid = (grid.abs().max(dim=-1)[0] <= 1).unsqueeze(1).repeat(1, 3, 1, 1)
    return warped, valid


def ssim(x, y):
    """networks/layers.py:148-178 -- 3x3 reflect-padded SSIM distance map, clamp((1-SSIM)/2, 0, 1)."""
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    x = F.pad(x, (1, 1, 1, 1), mode="reflect")
    y = F.pad(y, (1, 1, 1, 1), mode="reflect")
    mu_x = F.avg_pool2d(x, 3, 1)
    mu_y = F.avg_pool2d(y, 3, 1)
    sigma_x = F.avg_pool2d(x ** 2, 3, 1) - mu_x ** 2
    sigma_y = F.avg_pool2d(y ** 2, 3, 1) - mu_y ** 2
    sigma_xy = F.avg_pool2d(x * y, 3, 1) - mu_x * mu_y
    n = (2 * mu_x * mu_y + C1) * (2 * sigma_xy + C2)
    d = (mu_x ** 2 + mu_y ** 2 + C1) * (sigma_x + sigma_y + C2)
    return torch.clamp((1 - n / d) / 2, 0, 1)


def photo_metric_loss(target, reference, flow_map, pix_coords, use_ssim=True, padding_mode="zeros"):
    """loss_functions.py:107-115 -- (loss, warped, diff, valid)."""
    warped, valid = inverse_warp(reference, flow_map, pix_coords, padding_mode)
    diff = (target - warped).abs() * valid
    loss = diff.mean()
    if use_ssim:
        loss = 0.15 * loss + 0.85 * ssim(target, warped).mean()
    return loss, warped, diff, valid


# --------------------------------------------------------------------------- mask regularisers
def smooth_loss(target, mobile):
    """loss_utils.py:151-168 -- edge-aware first-order smoothness of the mobile map."""
    gix = torch.mean(torch.abs(target[:, :, :, :-1] - target[:, :, :, 1:]), 1, keepdim=True)
    giy = torch.mean(torch.abs(target[:, :, :-1, :] - target[:, :, 1:, :]), 1, keepdim=True)
    gmx = torch.abs(mobile[:, :, :, :-1] - mobile[:, :, :, 1:])
    gmy = torch.abs(mobile[:, :, :-1, :] - mobile[:, :, 1:, :])
    gmx *= torch.exp(-gix)
    gmy *= torch.exp(-giy)
    return gmx.mean() + gmy.mean()


def derivable_consistency_loss(mobile1, mobile2, threshold=0.5):
    """loss_utils.py:171-177."""
    a1 = torch.sigmoid(20 * (mobile1 - threshold))
    a2 = torch.sigmoid(20 * (mobile2 - threshold))
    return (a1 - a2) ** 2


def compute_quantiles(flow, cam_T_cam, inv_K, p1, pix_coords, ones, scale_factor, percentage, i, b):
    """loss_utils.py:197-202."""
    flow_map = scale_factor * flow[("flow", i, 0)]
    p2 = torch.cat([pix_coords + flow_map, ones], 1).view(b, 3, -1)
    e = get_epipolar_new(p1, p2, inv_K[:, :3, :3], cam_T_cam[:, :3, :3], cam_T_cam[:, :3, -1]).view(b, -1).abs()
    return torch.quantile(e, percentage, dim=1)


# --------------------------------------------------------------------------- per-pair epipolar loss
def epipolar_loss(flow_map, mobile_mask, instances_info, inv_K, ro, tran, pix_coords, *,
                  mode="DC", alpha=0.55, w_d2_sim=0.05, threshold=None, weight=None, ds_base="SN"):
    """loss_functions.py:117-138 with the section-8a-M mode switch.

    DC is HEAD (SN post + cross entropy, loss_functions.py:124,132-133).
    Returns (loss, post.expand(b,3,h,w), map.expand(b,3,h,w)) like the reference;
    in SN-based modes ``map`` is the NORMALISED map because post_process_epipolar_1
    divides in place (quirk kept).
    """
    assert mode in MODES
    b, _, h, w = flow_map.size()
    ones = torch.ones_like(mobile_mask)
    p1 = torch.cat([pix_coords, ones], 1).view(b, 3, -1)
    p2 = torch.cat([pix_coords + flow_map, ones], 1).view(b, 3, -1)
    emap = get_epipolar_new(p1, p2, inv_K[:, :3, :3], ro, tran).view(b, 1, h, w).abs()

    base = ds_base if mode == "DS" else ("SN" if mode in ("SN", "DC") else mode)
    if base == "SN":
        post = post_process_epipolar_1(emap)
    elif base == "T":
        post = post_pro_epipolar_weighted(emap, None, threshold)
    else:
        post = post_pro_epipolar_weighted(emap, weight, threshold)
    if mode == "DS":
        post = post_process_epipolar_2(post, instances_info)

    background = 1 - mobile_mask
    epipolar = (background * post).mean()
    non_trivial = (mobile_mask * torch.log(background + 1e-5)).abs().mean()
    loss = epipolar + alpha * non_trivial
    if mode == "DC":
        loss = epipolar + alpha * non_trivial + w_d2_sim * detectron2_similarity_loss(mobile_mask, instances_info).mean()
    return loss, post.expand(b, 3, h, w), emap.expand(b, 3, h, w)


# --------------------------------------------------------------------------- orchestration
class LossModule:
    """loss_functions.py:11-157, restated without nn.Module / device plumbing.

    Extra, non-upstream arguments (all keyword): ``mode``, ``weights`` (TG list),
    ``photometric`` (hooks `photo_metric_loss` back in exactly where the
    commented-out call sites are, loss_functions.py:48-50, with weight opt.w_p
    as in :194), ``ds_base``.
    """

    def __init__(self, opt, ssim_on=False, padding_mode="zeros", *, mode="DC", weights=None,
                 photometric=False, ds_base="SN"):
        self.opt, self.ssim_on, self.padding_mode = opt, ssim_on, padding_mode
        self.mode, self.weights, self.photometric, self.ds_base = mode, weights, photometric, ds_base
        self.losses = {"consis": 0, "epip": 0, "smooth": 0}
        if photometric:
            self.losses["photo"] = 0
        self.outputs = {"warps": {}, "diffs": {}, "valids": {}, "epipolars": {}, "flows": {}, "epipolar_ori": {}}

    def _epi(self, f, mobile, instances_info, inv_K, ro, tran, pix, scale):
        o = self.opt
        return epipolar_loss(f, mobile, instances_info, inv_K, ro, tran, pix, mode=self.mode, alpha=o.alpha,
                             w_d2_sim=o.w_d2_sim, threshold=getattr(o, "threshold", None),
                             weight=None if self.weights is None else self.weights[scale].to(f.device),
                             ds_base=self.ds_base)

    def frame_terms(self, inputs, i, flow, mobile, instances_info, cam_T_cam, scale):
        """Body of the per-frame loop, loss_functions.py:43-67 (== :88-105 for the single-mask form)."""
        tgt = inputs[("color", 0, scale)]
        b, _, h, w = tgt.size()
        pix = create_coords(b, h, w, tgt.device)
        avg = 2 ** scale
        f = get_scale_factor(b, h, w, tgt.device) * flow[("flow", i, scale)]
        ro, tran = cam_T_cam[i][:, :3, :3], cam_T_cam[i][:, :3, -1]
        if self.photometric:
            pl, warped, diff, valid = photo_metric_loss(tgt, inputs[("color", i, scale)], f, pix,
                                                        self.ssim_on, self.padding_mode)
            self.losses["photo"] = self.losses["photo"] + (pl / avg)
        if not self.opt.disable_smoothloss:
            self.losses["smooth"] = self.losses["smooth"] + (smooth_loss(tgt, mobile) / avg)
        el, emap, eori = self._epi(f, mobile, instances_info, inputs[("inv_K", scale)], ro, tran, pix, scale)
        self.losses["epip"] = self.losses["epip"] + (el / avg)
        if scale == 0:
            if self.photometric:
                self.outputs["warps"][(i, scale)] = warped
                self.outputs["diffs"][(i, scale)] = diff
                self.outputs["valids"][(i, scale)] = valid
            self.outputs["epipolars"][(i, scale)] = emap
            self.outputs["flows"][(i, scale)] = f
            self.outputs["epipolar_ori"][(i, scale)] = eori

    def consistency_loss(self, m1, m2, scale):
        """loss_functions.py:140-147."""
        self.losses["consis"] = self.losses["consis"] + derivable_consistency_loss(m1, m2).mean() / (2 ** scale)


def loss_forward(opt, inputs, frame_ids, flow, mobile, instances_info, scales, cam_T_cam, *, ssim_on=False,
                 padding_mode="zeros", mode="DC", weights=None, photometric=False, ds_base="SN"):
    """Loss.forward, loss_functions.py:170-205 -> (outputs, losses)."""
    lm = LossModule(opt, ssim_on, padding_mode, mode=mode, weights=weights, photometric=photometric, ds_base=ds_base)
    min_mobiles = {}
    for s in scales:
        m1, m2 = mobile[("mobile", -1, s)], mobile[("mobile", 1, s)]
        min_mobiles[s] = torch.cat([m1, m2], dim=1).min(1, True)[0]
        if not opt.disable_consisloss:
            lm.consistency_loss(m1, m2, s)
        for i in frame_ids:
            m = mobile[("mobile", i, s)] if opt.disable_min else min_mobiles[s]
            lm.frame_terms(inputs, i, flow, m, instances_info, cam_T_cam, s)
    losses = lm.losses
    losses["loss"] = opt.w_e * losses["epip"] + opt.w_s * losses["smooth"] + opt.w_c * losses["consis"]
    if photometric:
        losses["loss"] = losses["loss"] + opt.w_p * losses["photo"]
    outputs = lm.outputs
    outputs["min_mobiles"] = min_mobiles
    return outputs, losses
