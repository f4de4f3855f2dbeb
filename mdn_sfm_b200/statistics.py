"""Dataset-wide epipolar statistics: the thresholds of the T / TG modes.

Mirrors ``Trainer.epipolar_statics`` (trainer.py:521-562): for every batch and source frame, |e| of the pixel grid under
the predicted flow and pose, 1000 per-sample quantiles, and at the end the percentiles of all of them (80 ... 99) -- the
numbers ``options.py:84-87`` / ``options_eval.py:55-58`` hard-code for 128x416 and which have to be re-derived for any
other resolution.  HEAD calls ``get_epipolar_new`` there with the wrong arity (it would raise); this is the working form,
equal to ``loss_utils.compute_quantiles`` (:197-202) per batch.

The per-pixel map comes from the fused kernel's maps-only launch (``ori_map`` in T mode without a threshold is |e|), the
quantiles from ``torch.quantile`` on the device; nothing leaves the GPU until the final percentiles.  ``weight`` divides
|e| first (the TG statistic, the commented-out line trainer.py:553).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _cabi, fused
from ._cabi import MASK_SHARED, TERM_EPIPOLAR

PERCENTILES = (80, 85, 88, 90, 92, 95, 98, 99)


def epipolar_abs_map(flow, inv_K, cam_T_cam, library=None, arith="cuda"):
    """|e| (B,1,h,w) of the pixel grid: p1 = (x, y, 1), p2 = p1 + [w, h] * flow (loss_functions.py:44,120-123), e from
    get_epipolar_new (loss_utils.py:39-69).  flow (B,2,h,w) in the nets' normalised units, inv_K / cam_T_cam (B,4,4)."""
    flow = _cabi.check_tensor(flow, what="flow").contiguous()
    B, _, h, w = flow.shape
    S = fused.ScaleData(h, w, float(w), float(h), 1.0)
    S.flow[0] = flow.detach()
    S.mob[0] = torch.zeros((B, 1, h, w), dtype=torch.float32, device=flow.device)     # bg = 1: the map is unmasked anyway
    cfg = fused.FusedConfig(cuda_arith=arith == "cuda", batch=B, n_pairs=1, post=fused.POST_T, mask_mode=MASK_SHARED,
                            flags=TERM_EPIPOLAR, threshold=None, want_maps=("ori_map",))
    if arith == "cuda":
        with torch.no_grad():
            _, _, maps = fused.fused_loss(cfg, [S], library, cams=[cam_T_cam.detach().contiguous()],
                                          inv_Ks=[inv_K.detach().contiguous()])
    else:
        S.fmat[0] = fused.fundamental_matrix(inv_K[:, :3, :3], cam_T_cam[:, :3, :3], cam_T_cam[:, :3, -1]).detach().contiguous()
        with torch.no_grad():
            _, _, maps = fused.fused_loss(cfg, [S], library)
    return maps["ori_map"][0]


class EpipolarStatistics:
    """Accumulates per-sample quantiles of |e| (optionally |e| / weight) over a data set, per source frame."""

    def __init__(self, frame_ids=(-1, 1), num_quantile=1000, weight=None, library=None, arith="cuda"):
        self.frame_ids, self.num_quantile = tuple(frame_ids), num_quantile
        self.weight, self.library, self.arith = weight, library, arith
        self.percentiles = {i: [] for i in self.frame_ids}
        self._q = None

    def update(self, flows, inv_K, cam_T_cams):
        """flows: {("flow", i, 0): (B,2,h,w)}, inv_K (B,4,4) of scale 0, cam_T_cams: {i: (B,4,4)} (trainer.py:541-555)."""
        for i in self.frame_ids:
            e = epipolar_abs_map(flows[("flow", i, 0)], inv_K, cam_T_cams[i], self.library, self.arith)
            if self.weight is not None:
                e = e / self.weight.to(e.device)
            if self._q is None or self._q.device != e.device:
                self._q = torch.linspace(0, 1, self.num_quantile, device=e.device)
            self.percentiles[i].append(torch.quantile(e.view(e.shape[0], -1), self._q, dim=1))     # (Q, B)

    def result(self):
        """-> (percentiles (n_frames, Q, total samples) numpy, thresholds at PERCENTILES) like trainer.py:557-562."""
        per = torch.stack([torch.cat(self.percentiles[i], 1) for i in self.frame_ids], 0).cpu().numpy()
        return per, np.percentile(per.reshape(-1), PERCENTILES)
