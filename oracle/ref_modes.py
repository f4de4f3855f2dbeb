"""Mode compositions built from the REFERENCE's own functions (build container only).

The upstream tree has no runtime mode switch: HEAD is
"DC", and SN / T / TG / DS exist as functions plus commented-out call sites.
This module defines each mode as a temporary method patch of the reference ``LossModule`` whose
``epipolar_loss`` differs from loss_functions.py:117-138 only in the lines the
upstream comments toggle, always calling the reference's functions.  It is what
``oracle/make_golden.py`` runs to produce fixtures and what
``tests/test_oracle_vs_reference.py`` checks ``oracle/restate.py`` against.
TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import torch

from . import ref_loader


def mode_loss_module(mode, weights=None, photometric=False, ds_base="SN"):
    ref = ref_loader.load()
    lu = ref.loss_utils
    Base = ref_loader.force_cpu_loss_module()
    base_forward, base_single = Base.forward, Base.single_mobile_mask_forward

    class ModeMethods:
        def epipolar_loss(self, flow_map, mobile_mask, instances_info, inv_K, ro, tran):
            b, _, h, w = flow_map.size()
            ones = torch.ones_like(mobile_mask)
            p1 = torch.cat([self.pix_coords, ones], 1).view(b, 3, -1)
            p2 = torch.cat([self.pix_coords + flow_map, ones], 1).view(b, 3, -1)
            emap = lu.get_epipolar_new(p1, p2, inv_K[:, :3, :3], ro, tran).view(b, 1, h, w).abs()
            base = ds_base if mode == "DS" else ("SN" if mode in ("SN", "DC") else mode)
            if base == "SN":
                post = lu.post_process_epipolar_1(emap)
            else:
                wgt = None
                if base == "TG":
                    scale = {weights[k].shape[-1]: k for k in range(len(weights))}[w]
                    wgt = weights[scale]
                post = lu.post_pro_epipolar_weighted(emap, wgt, self.options.threshold)
            if mode == "DS":
                post = lu.post_process_epipolar_2(post, instances_info)
            background = 1 - mobile_mask
            epipolar = (background * post).mean()
            non_trivial = (mobile_mask * torch.log(background + 1e-5)).abs().mean()
            loss = epipolar + self.alpha * non_trivial
            if mode == "DC":
                ce = lu.detectron2_similarity_loss(mobile_mask, instances_info).mean()
                loss = epipolar + self.alpha * non_trivial + self.options.w_d2_sim * ce
            return loss, post.expand(b, 3, h, w), emap.expand(b, 3, h, w)

        def _photo_terms(self, inputs, frame_ids, flow, scale):
            # re-enable the commented-out call sites loss_functions.py:48-50 / :89-91 / :194
            tgt = inputs[("color", 0, scale)]
            b, _, h, w = tgt.size()
            self.pix_coords = self.create_coords(b, h, w)
            sf = ref.layers.get_scale_factor(b, h, w)
            for i in frame_ids:
                f = sf * flow[("flow", i, scale)]
                pl, warped, diff, valid = self.photo_metric_loss(tgt, inputs[("color", i, scale)], f)
                self.losses["photo"] = self.losses.get("photo", 0) + pl / (2 ** scale)
                if scale == 0:
                    self.outputs["warps"][(i, scale)] = warped
                    self.outputs["diffs"][(i, scale)] = diff
                    self.outputs["valids"][(i, scale)] = valid

        def forward(self, inputs, frame_ids, flow, mobile, instances_info, cam_T_cam, scale):
            if photometric:
                ModeMethods._photo_terms(self, inputs, frame_ids, flow, scale)
            return base_forward(self, inputs, frame_ids, flow, mobile, instances_info, cam_T_cam, scale)

        def single_mobile_mask_forward(self, inputs, frame_id, flow, mobile, instances_info, cam_T_cam, scale):
            if photometric:
                ModeMethods._photo_terms(self, inputs, [frame_id], flow, scale)
            return base_single(self, inputs, frame_id, flow, mobile, instances_info, cam_T_cam, scale)

    return ModeMethods


def reference_loss_forward(opt, inputs, frame_ids, flow, mobile, instances_info, scales, cam_T_cam, *,
                           mode="DC", weights=None, photometric=False, ssim_on=False, ds_base="SN"):
    """Run the reference ``Loss.forward`` (loss_functions.py:170-205) with LossModule swapped for the mode class."""
    ref = ref_loader.load()
    lf = ref.loss_functions
    methods = mode_loss_module(mode, weights, photometric, ds_base)
    cls = lf.LossModule
    saved = (cls.epipolar_loss, cls.forward, cls.single_mobile_mask_forward)
    cls.epipolar_loss, cls.forward = methods.epipolar_loss, methods.forward
    cls.single_mobile_mask_forward = methods.single_mobile_mask_forward
    try:
        loss = lf.Loss(opt, no_ssim=not ssim_on)
        if not ssim_on:
            loss.ssim = None  # upstream bug: no_ssim=True leaves self.ssim undefined (loss_functions.py:163-164,172)
        outputs, losses = loss(inputs, list(frame_ids), flow, mobile, instances_info, list(scales), cam_T_cam)
    finally:
        cls.epipolar_loss, cls.forward, cls.single_mobile_mask_forward = saved
    if photometric:
        losses["loss"] = losses["loss"] + opt.w_p * losses["photo"]
    return outputs, losses
