"""Drop-in for the loss-path helpers of the reference ``networks/layers.py``."""
from __future__ import annotations

import torch
from torch import nn

from .ops import SsimFn, _c


def get_scale_factor(batch_size, height, width):
    """networks/layers.py:101-103 -> [W,H] broadcast to (B,2,H,W) (a stride-0 view; the kernels take two scalars)."""
    return torch.tensor([float(width), float(height)]).view(1, 2, 1, 1).expand(batch_size, 2, height, width)


def rot_from_axisangle(vec):
    """networks/layers.py:59-98 -> (B,4,4).  Closed form (Rodrigues) built with a handful of batched ops
    instead of ~25 scalar-indexed writes; same formula, same 1e-7 guard."""
    angle = torch.norm(vec, 2, 2, True)
    axis = vec / (angle + 1e-7)
    ca, sa = torch.cos(angle), torch.sin(angle)
    C = 1 - ca
    x, y, z = axis[..., 0:1], axis[..., 1:2], axis[..., 2:3]
    xs, ys, zs = x * sa, y * sa, z * sa
    xC, yC, zC = x * C, y * C, z * C
    xyC, yzC, zxC = x * yC, y * zC, z * xC
    zero, one = torch.zeros_like(ca), torch.ones_like(ca)
    rows = [x * xC + ca, xyC - zs, zxC + ys, zero,
            xyC + zs, y * yC + ca, yzC - xs, zero,
            zxC - ys, yzC + xs, z * zC + ca, zero,
            zero, zero, zero, one]
    return torch.cat(rows, 2).view(vec.shape[0], 4, 4)


def get_translation_matrix(translation_vector):
    """networks/layers.py:43-56."""
    B = translation_vector.shape[0]
    T = torch.eye(4, device=translation_vector.device, dtype=translation_vector.dtype).repeat(B, 1, 1)
    return torch.cat([T[:, :, :3], torch.cat([translation_vector.contiguous().view(B, 3, 1), T[:, 3:, 3:]], 1)], 2)


class PoseParameters:
    """PoseNet's outputs for one source frame -- axisangle, translation, each (B,1,1,3) -- handed to ``Loss.forward`` AS
    PARAMETERS in place of the (B,4,4) matrix (``cam_T_cam[i] = PoseParameters(axisangle, translation)``).  The fused call
    then runs transformation_from_parameters itself (MdnLossDesc.axisangle / translation) and returns the gradients
    w.r.t. both: the ~40 small launches of networks/layers.py:16-98 and their autograd backward leave the training step.
    ``matrix()`` materialises the reference's (B,4,4) tensor for any other consumer."""

    def __init__(self, axisangle, translation):
        if tuple(axisangle.shape[1:]) != (1, 1, 3) or tuple(translation.shape) != tuple(axisangle.shape):
            raise ValueError("axisangle / translation must be (B,1,1,3) (pose_net_v3.py:62-64)")
        self.axisangle, self.translation = axisangle, translation

    def matrix(self, invert=False):
        return transformation_from_parameters(self.axisangle, self.translation, invert)


def transformation_from_parameters(axis_angle, translation, invert=False):
    """networks/layers.py:16-40."""
    R = rot_from_axisangle(axis_angle.squeeze(1))
    t = translation.clone().squeeze(1)
    if invert:
        R = R.transpose(1, 2)
        t = t * -1
    T = get_translation_matrix(t)
    return torch.matmul(R, T) if invert else torch.matmul(T, R)


class SSIM(nn.Module):
    """networks/layers.py:148-178: clamp((1 - SSIM(x, y)) / 2, 0, 1), 3x3 windows, reflect padding."""

    def __init__(self, library=None):
        super().__init__()
        self.C1, self.C2 = 0.01 ** 2, 0.03 ** 2
        self._library = library

    def forward(self, x, y):
        return SsimFn.apply(_c(x, "x"), _c(y, "y"), self._library)
