"""Synthetic KITTI-shaped inputs for the loss path.

There is no KITTI data, no trained FlowNet/PoseNet/MDN and no Detectron2 in the
image, so every test, ``smoke()`` and ``bench.py`` run on tensors of the right
shape and statistics, generated on the CPU from a seeded ``torch.Generator`` so
they are identical in the build container and on the GPU box:

* images: U[0,1) then ``(x - 0.45) / 0.225`` (mono_dataset.py:51-52,112),
  lower scales by torchvision ``Resize`` (mono_dataset.py:122-125);
* intrinsics: KITTI's normalised K scaled to the image (kitti_dataset.py:35-38),
  rows 0,1 divided by ``2**s``, ``inv_K = pinv(K)`` (mono_dataset.py:113-121);
* flow: N(0, flow_std^2) in the nets' normalised units (fraction of the image);
* mobile masks: U(0.02, 0.98); pose: axis-angle / translation N(0, std^2) run
  through ``transformation_from_parameters``;
* Detectron2 stand-in: objects with a bool ``pred_masks`` (N,375,1242).
"""
from __future__ import annotations

from types import SimpleNamespace

import torch

INSTANCE_HW = (375, 1242)  # the size the reference feeds Detectron2 (mono_dataset.py:111)


class SyntheticInstances:
    """Stands in for ``detectron2.structures.Instances``: only ``pred_masks`` is read (loss_utils.py:113,118)."""

    def __init__(self, pred_masks):
        self.pred_masks = pred_masks

    def to(self, device):
        return SyntheticInstances(self.pred_masks.to(device))


def default_opt(batch_size, height, width, **over):
    """The ``opt`` fields the loss reads (options.py:56-109,145-170) with the reference defaults."""
    o = dict(alpha=0.55, threshold=9.22, batch_size=batch_size, height=height, width=width,
             disable_smoothloss=False, disable_consisloss=False, disable_min=False, disable_photoloss=False,
             w_e=1.0, w_s=1.0, w_c=0.5, w_p=1.0, w_d2_sim=0.05, no_ssim=False,
             scales=[0, 1, 2, 3], frame_ids=[0, -1, 1])
    o.update(over)
    return SimpleNamespace(**o)


def _rot_from_axisangle(vec):
    # Same closed form as networks/layers.py:59-98; used here only to manufacture inputs.
    angle = vec.norm(dim=2, keepdim=True)
    axis = vec / (angle + 1e-7)
    ca, sa = torch.cos(angle), torch.sin(angle)
    C = 1 - ca
    x, y, z = axis[..., 0:1], axis[..., 1:2], axis[..., 2:3]
    rows = [torch.cat([x * x * C + ca, x * y * C - z * sa, z * x * C + y * sa], 2),
            torch.cat([x * y * C + z * sa, y * y * C + ca, y * z * C - x * sa], 2),
            torch.cat([z * x * C - y * sa, y * z * C + x * sa, z * z * C + ca], 2)]
    return torch.cat(rows, 1)  # (B,3,3)


def make_pose(axis_angle, translation):
    """(B,1,1,3) x2 -> (B,4,4) with M[:3,:3]=R, M[:3,3]=t (the invert=False form, trainer.py:272)."""
    B = axis_angle.shape[0]
    M = torch.zeros(B, 4, 4)
    M[:, :3, :3] = _rot_from_axisangle(axis_angle.reshape(B, 1, 3))
    M[:, :3, 3] = translation.reshape(B, 3)
    M[:, 3, 3] = 1
    return M


def make_instances(batch, gen, n_inst=3):
    out = []
    H, W = INSTANCE_HW
    for _ in range(batch):
        masks = torch.rand(n_inst, H, W, generator=gen) > 0.9
        for n in range(n_inst):
            y0 = int(torch.randint(0, H - 60, (1,), generator=gen))
            x0 = int(torch.randint(0, W - 200, (1,), generator=gen))
            hh = int(torch.randint(20, 60, (1,), generator=gen))
            ww = int(torch.randint(40, 200, (1,), generator=gen))
            masks[n, y0:y0 + hh, x0:x0 + ww] = True
        out.append({"instances": SyntheticInstances(masks)})
    return out


def smooth_flow(batch, h, w, flow_std, gen, cell=32):
    """A network-like flow field: what a (random-init or trained) FlowNet decoder emits is spatially smooth, not
    white noise.  Per-sample global motion + a smooth field (white noise on a coarse grid, one node per ``cell``
    full-resolution pixels, bilinearly upsampled like the decoder's own upsampling) + 5 % pixel-level jitter; the
    total standard deviation is ``flow_std`` in the nets' normalised units."""
    import torch.nn.functional as F
    gh, gw = max(2, -(-h // cell) + 1), max(2, -(-w // cell) + 1)
    coarse = torch.randn(batch, 2, gh, gw, generator=gen)
    field = F.interpolate(coarse, size=(h, w), mode="bilinear", align_corners=True)
    glob = torch.randn(batch, 2, 1, 1, generator=gen)
    jitter = torch.randn(batch, 2, h, w, generator=gen)
    return flow_std * (0.6 * glob + 0.8 * field + 0.05 * jitter)


def make_batch(batch, height, width, scales=(0, 1, 2, 3), frame_ids=(-1, 1), seed=42, flow_std=0.01,
               pose_rot_std=0.01, pose_t_std=0.1, with_instances=True, device="cpu", flow_kind="iid"):
    """Returns ``(inputs, flows, mobiles, cam_T_cam, instances)`` laid out as trainer.py:256-281 hands them on.

    ``flow_kind``: "iid" = white noise N(0, flow_std^2) per pixel (the stress case: every pixel gathers from an
    unrelated place), "smooth" = a network-like field of the same magnitude (``smooth_flow``)."""
    from torchvision.transforms import Resize

    gen = torch.Generator().manual_seed(seed)
    inputs, flows, mobiles, cams = {}, {}, {}, {}
    K = torch.tensor([[0.58 * width, 0, 0.5 * width, 0], [0, 1.92 * height, 0.5 * height, 0],
                      [0, 0, 1, 0], [0, 0, 0, 1]], dtype=torch.float32)
    for i in (0,) + tuple(frame_ids):
        img = (torch.rand(batch, 3, height, width, generator=gen) - 0.45) / 0.225
        for s in scales:
            inputs[("color", i, s)] = img if s == 0 else Resize((height // 2 ** s, width // 2 ** s))(img)
    for s in scales:
        Ks = K.clone()
        Ks[0] /= 2 ** s
        Ks[1] /= 2 ** s
        inputs[("inv_K", s)] = torch.linalg.pinv(Ks).unsqueeze(0).repeat(batch, 1, 1)
    for i in frame_ids:
        for s in scales:
            h, w = height // 2 ** s, width // 2 ** s
            if flow_kind == "smooth":
                flows[("flow", i, s)] = smooth_flow(batch, h, w, flow_std, gen, cell=max(2, 32 >> s))
            else:
                flows[("flow", i, s)] = torch.randn(batch, 2, h, w, generator=gen) * flow_std
            mobiles[("mobile", i, s)] = torch.rand(batch, 1, h, w, generator=gen) * 0.96 + 0.02
        aa = torch.randn(batch, 1, 1, 3, generator=gen) * pose_rot_std
        tt = torch.randn(batch, 1, 1, 3, generator=gen) * pose_t_std
        cams[i] = make_pose(aa, tt)
    instances = make_instances(batch, gen) if with_instances else None

    def mv(d):
        return {k: v.to(device) for k, v in d.items()}

    if instances is not None and str(device) != "cpu":
        instances = [{"instances": d["instances"].to(device)} for d in instances]
    return mv(inputs), mv(flows), mv(mobiles), mv(cams), instances
