"""Drop-in for the loss-side helpers of the reference ``utils.py`` (binary_image, FlowWarp, gauss_distance_weight)."""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from . import _cabi
from .ops import FlowWarpFn, _c


def binary_image(x, threshold=0.5, library=None):
    """utils.py:100-103: 1.0 where x >= threshold else 0."""
    x = _c(x, "x")
    out = torch.empty_like(x)
    (library or _cabi.lib()).call("mdn_binary_image", x.data_ptr(), out.data_ptr(), x.numel(), float(threshold),
                                  _cabi.stream_ptr(x), dev=x)
    return out


class FlowWarp(nn.Module):
    """utils.py:289-315: forward(flow) -> (pix_coords, pix_coords_norm, valid_points), no sampling."""

    def __init__(self, batch_size, height, width, library=None, arith="cuda"):
        super().__init__()
        self.batch_size, self.height, self.width = batch_size, height, width
        self._library = library
        self._flags = _cabi.WARP_FLOWWARP_NORM | (_cabi.WARP_CUDA_ARITH if arith == "cuda" else 0)

    def forward(self, flow):
        flow = _c(flow, "flow")
        B, _, h, w = flow.shape
        _, grid, valid = FlowWarpFn.apply(flow, flow, self._flags, False, self._library)
        xs = torch.arange(w, dtype=torch.float32, device=flow.device).view(1, 1, 1, w)
        ys = torch.arange(h, dtype=torch.float32, device=flow.device).view(1, 1, h, 1)
        pix = torch.cat([xs + flow[:, 0:1], ys + flow[:, 1:2]], 1)
        return pix, grid, valid.bool()


_gauss_cache = {}


def gauss_distance_weight(num_scale, height=128, width=416, sigma1=30, sigma2=120):
    """utils.py:355-379: TG weights, one (1,1,h,w) fp32 tensor per scale, computed in float64.

    W[i,j] = 2e5 * (G.max() - G[i,j]) + 5 with G a centred 2-D Gaussian (rho = 0).  Vectorised instead of the
    reference's Python double loop (0.7 s at 192x640); the expression order inside the exponent is kept so the
    float64 values agree before the cast to fp32.  Cached per argument tuple (the table only depends on shape).
    """
    key = (num_scale, height, width, sigma1, sigma2)
    if key not in _gauss_cache:
        out = []
        for s in range(num_scale):
            num = 2 ** s
            h, w = height // num, width // num
            i = np.arange(h, dtype=np.float64).reshape(h, 1)
            j = np.arange(w, dtype=np.float64).reshape(1, w)
            a = (i - h // 2) ** 2 / (sigma1 / num) ** 2
            b = (j - w // 2) ** 2 / (sigma2 / num) ** 2
            factor = num ** 2 / (2 * np.pi * sigma1 * sigma2 * 1.0) / num ** 2
            g = factor * np.exp(-(a + b - 0.0) / 2.0)
            out.append(torch.tensor(2e5 * (g.max() - g) + 5).unsqueeze(0).unsqueeze(0).type(torch.float32))
        _gauss_cache[key] = out
    return [t.clone() for t in _gauss_cache[key]]
