"""Host side of the fused loss call: tensor marshalling + the autograd node.

One ``mdn_loss_fused`` call evaluates every requested term for all scales and both source frames and, when any
input requires grad, also writes the gradients for an upstream gradient of 1 (the loss is a scalar, so the
backward pass is a rescale -- ``mdn_loss_scale_grads`` -- that returns immediately for ``loss.backward()``).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import torch

from . import _cabi
from ._cabi import (MASK_MIN, MASK_OWN, MASK_SHARED, OPT_CROSS_ENT, OPT_CUDA_ARITH, OPT_GRADS, OPT_INST_MASK, OPT_SSIM, OUT_APPLIED,
                    OUT_COUNT, POST_SN, POST_T, POST_TG, TERM_CONSIS, TERM_EPIPOLAR, TERM_PHOTO, TERM_SMOOTH)

POST_OF_MODE = {"SN": POST_SN, "T": POST_T, "TG": POST_TG}


def fundamental_matrix(inv_K, rotation, translation):
    """F = K^-T ((t_x R) K^-1), the three small matmuls of loss_utils.py:50-62 in the reference's association order.

    Stays in torch on the host side of the boundary: parity of F is then exact by
    construction and autograd carries the kernel's d(loss)/dF back to rotation / translation / PoseNet.
    Broadcasts: inv_K (..., B, 3, 3), rotation (..., B, 3, 3), translation (..., B, 3).
    """
    t0, t1, t2 = translation[..., 0], translation[..., 1], translation[..., 2]
    z = torch.zeros_like(t0)
    t_x = torch.stack([z, -t2, t1, t2, z, -t0, -t1, t0, z], dim=-1).reshape(translation.shape[:-1] + (3, 3))
    return torch.matmul(inv_K.transpose(-2, -1), torch.matmul(torch.matmul(t_x, rotation), inv_K))


@dataclass
class ScaleData:
    """Tensors of one pyramid level, in the order the C ABI wants them."""
    height: int
    width: int
    flow_sx: float
    flow_sy: float
    scale_div: float
    tgt: Optional[torch.Tensor] = None
    ref: List[Optional[torch.Tensor]] = field(default_factory=lambda: [None, None])
    ref_packed: List[Optional[torch.Tensor]] = field(default_factory=lambda: [None, None])   # (B,h,w,4): no repack kernel
    flow: List[Optional[torch.Tensor]] = field(default_factory=lambda: [None, None])
    mob: List[Optional[torch.Tensor]] = field(default_factory=lambda: [None, None])
    fmat: List[Optional[torch.Tensor]] = field(default_factory=lambda: [None, None])
    weight: Optional[torch.Tensor] = None
    inst: Optional[torch.Tensor] = None


@dataclass
class FusedConfig:
    batch: int
    n_pairs: int
    post: int
    mask_mode: int
    flags: int
    threshold: Optional[float] = None
    alpha: float = 0.0
    w_d2_sim: float = 0.0
    w_e: float = 1.0
    w_s: float = 1.0
    w_c: float = 1.0
    w_p: float = 1.0
    cuda_arith: bool = True   # replay the reference's CUDA-eager rounding (MDN_OPT_CUDA_ARITH); False = its CPU rounding
    want_maps: tuple = ()   # subset of ("post_map", "ori_map", "warped", "diff", "valid", "ssim_map"), scale 0 only
    inst_ready: object = None   # torch.cuda.Event recorded on another stream after the instance-mask preparation (MdnLossDesc.inst_ready)


_workspaces = {}


def _workspace(device, nbytes):
    key = (str(device), torch.cuda.current_stream(device).cuda_stream if device.type == "cuda" else 0)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)   # needs no initialisation
        _workspaces[key] = ws
    return ws


def run_fused(cfg: FusedConfig, scales: List[ScaleData], need_grad, library=None, g_fmat_all=None, poses=None, g_cams=None,
              pose_params=None, g_pose_params=None):
    """Launches the fused kernel.

    need_grad: per scale a dict {"flow": [bool,bool], "mob": [bool,bool], "fmat": [bool,bool]}.
    g_fmat_all: optional (S,P,B,3,3) buffer whose [k,p] slices receive d/dF (one tensor for the whole pyramid).
    poses: optional (cams, inv_Ks) -- one (B,4,4) pose per pair and one (B,4,4) inverse intrinsics per scale; the
    kernels then build the fundamental matrices themselves (MdnLossDesc.cam / inv_K) and `fmat` entries are ignored.
    g_cams: with poses, per pair a (B,4,4) buffer (or None) that receives d(loss)/d(pose).
    pose_params: optional (axisangles, translations, inv_Ks) INSTEAD of poses -- PoseNet's (B,1,1,3) outputs per pair; the
    kernels build the pose matrices too (transformation_from_parameters, networks/layers.py:16-98).
    g_pose_params: with pose_params, (g_axisangles, g_translations): per pair a (B,1,1,3) buffer or None.
    Returns (loss_out (8,), grads (same structure, tensors or None), maps (dict name -> [pair tensors]), call).
    """
    library = library or _cabi.lib()
    dev = None
    any_grad = (any(any(v) for ng in need_grad for v in ng.values()) or g_fmat_all is not None
                or (g_cams is not None and any(g is not None for g in g_cams))
                or (g_pose_params is not None and any(g is not None for seq in g_pose_params for g in seq)))
    flags = cfg.flags | (OPT_GRADS if any_grad else 0) | (OPT_CUDA_ARITH if cfg.cuda_arith else 0)
    call = _cabi.FusedCall(batch=cfg.batch, n_pairs=cfg.n_pairs, post=cfg.post, mask_mode=cfg.mask_mode, flags=flags,
                           threshold=cfg.threshold, alpha=cfg.alpha, w_d2_sim=cfg.w_d2_sim, w_e=cfg.w_e, w_s=cfg.w_s,
                           w_c=cfg.w_c, w_p=cfg.w_p)
    grads, maps = [], {}
    B = cfg.batch
    for k, (S, ng) in enumerate(zip(scales, need_grad)):
        some = S.flow[0] if S.flow[0] is not None else (S.mob[0] if S.mob[0] is not None else S.tgt)
        dev = some.device
        h, w = S.height, S.width
        g = {"flow": [None, None], "mob": [None, None], "fmat": [None, None]}
        for p in range(2):
            if ng["flow"][p]:
                g["flow"][p] = torch.empty((B, 2, h, w), dtype=torch.float32, device=dev)
            if ng["mob"][p]:
                g["mob"][p] = torch.empty((B, 1, h, w), dtype=torch.float32, device=dev)
            if g_fmat_all is not None and p < cfg.n_pairs:
                g["fmat"][p] = g_fmat_all[k, p]
            elif ng["fmat"][p]:
                g["fmat"][p] = torch.empty((B, 3, 3), dtype=torch.float32, device=dev)
        if any_grad and cfg.mask_mode == MASK_MIN and (g["mob"][0] is None) != (g["mob"][1] is None):
            # the arg-min routing writes both maps; give the kernel a scratch target for the unused one
            for p in range(2):
                if g["mob"][p] is None:
                    g["mob"][p] = torch.empty((B, 1, h, w), dtype=torch.float32, device=dev)
        extra = {}
        if k == 0:
            for name in cfg.want_maps:
                if name == "valid":
                    bufs = [torch.empty((B, 1, h, w), dtype=torch.uint8, device=dev) for _ in range(cfg.n_pairs)]
                elif name in ("post_map", "ori_map"):
                    bufs = [torch.empty((B, 1, h, w), dtype=torch.float32, device=dev) for _ in range(cfg.n_pairs)]
                else:
                    bufs = [torch.empty((B, 3, h, w), dtype=torch.float32, device=dev) for _ in range(cfg.n_pairs)]
                maps[name] = bufs
                extra[name] = bufs
        call.add_scale(h, w, S.flow_sx, S.flow_sy, S.scale_div, tgt=S.tgt, ref=S.ref, ref_packed=S.ref_packed, flow=S.flow, mob=S.mob,
                       fmat=S.fmat, weight=S.weight, inst=S.inst, g_flow=g["flow"], g_mob=g["mob"], g_fmat=g["fmat"],
                       **extra)
        grads.append(g)
    if poses is not None:
        call.set_poses(poses[0], poses[1], g_cams)
    if pose_params is not None:
        gp = g_pose_params or (None, None)
        call.set_pose_params(pose_params[0], pose_params[1], pose_params[2], gp[0], gp[1])
    if cfg.inst_ready is not None:
        call.desc.inst_ready = cfg.inst_ready.cuda_event
        call.keep.append(cfg.inst_ready)
    loss_out = torch.empty(OUT_COUNT, dtype=torch.float32, device=dev)
    ws = _workspace(dev, call.workspace_bytes(library))
    call.run(library, loss_out, ws, _cabi.stream_ptr(loss_out))
    return loss_out, grads, maps, call


class _FusedLossFn(torch.autograd.Function):
    """total = fused(flows, mobs, fmats); the gradients were produced by the forward launch."""

    @staticmethod
    def forward(ctx, cfg, scales, slots, library, poses, *diff_inputs):
        # slots[i] = (scale index, kind, pair) for diff_inputs[i]; poses = ("cam", cams, inv_Ks) | ("params", aas, trs, inv_Ks) | None
        need = [{"flow": [False, False], "mob": [False, False], "fmat": [False, False]} for _ in scales]
        g_fmat_all = None
        mats = poses[1:] if poses is not None and poses[0] == "cam" else None
        params = poses[1:] if poses is not None and poses[0] == "params" else None
        g_cams = [None, None] if mats is not None else None
        g_par = ([None, None], [None, None]) if params is not None else None
        for (k, kind, p), t in zip(slots, diff_inputs):
            if kind == "fmat_all":
                g_fmat_all = torch.empty_like(t)
            elif kind == "cam":
                g_cams[p] = torch.empty_like(t)
            elif kind in ("aa", "tr"):
                g_par[0 if kind == "aa" else 1][p] = torch.empty(t.shape, dtype=torch.float32, device=t.device)
            elif t.requires_grad:
                need[k][kind][p] = True
        loss_out, grads, maps, call = run_fused(cfg, scales, need, library, g_fmat_all, mats, g_cams, params, g_par)
        ctx.call, ctx.library, ctx.slots, ctx.grads, ctx.loss_out = call, library, slots, grads, loss_out
        ctx.g_fmat_all, ctx.g_cams, ctx.g_par = g_fmat_all, g_cams, g_par
        ctx.maps = maps
        total = loss_out[0]
        terms = loss_out[1:5]
        ctx.mark_non_differentiable(terms)
        ctx.set_materialize_grads(False)   # no zero-filled gradient tensor (a fill launch) for `terms`
        return total, terms

    @staticmethod
    def backward(ctx, g_total, _g_terms):
        g = g_total.contiguous()
        if g.dtype != torch.float32:
            g = g.float()
        if ctx.grads is None:
            raise RuntimeError("mdn_sfm_b200: the fused loss hands its gradient buffers to autograd in backward(); a second "
                               "backward through the same forward (retain_graph=True) is not supported -- run the forward again")
        ctx.call.scale_grads(ctx.library, g, ctx.loss_out[OUT_APPLIED:], _cabi.stream_ptr(g))
        # Give the buffers AWAY: with no other reference left, AccumulateGrad adopts them as `.grad` instead of cloning
        # them (18 device-to-device copies, 94 MB of traffic and 45 us per step at the headline shape otherwise).
        grads, g_fmat_all, g_cams, g_par = ctx.grads, ctx.g_fmat_all, ctx.g_cams, ctx.g_par
        ctx.grads = ctx.g_fmat_all = ctx.g_cams = ctx.g_par = None
        ctx.call.keep = []     # (stream-ordered allocator: the launches above keep using the memory safely)
        out = [None, None, None, None, None]
        for (k, kind, p) in ctx.slots:
            if kind == "fmat_all":
                out.append(g_fmat_all)
            elif kind == "cam":
                out.append(g_cams[p])
            elif kind in ("aa", "tr"):
                out.append(g_par[0 if kind == "aa" else 1][p])
            else:
                out.append(grads[k][kind][p])
        return tuple(out)


def fused_loss(cfg: FusedConfig, scales: List[ScaleData], library=None, fmat_all=None, cams=None, inv_Ks=None, axisangles=None,
               translations=None):
    """-> (total 0-d tensor with grad_fn, terms (4,) = [epip, smooth, consis, photo] detached, maps dict).

    fmat_all: the (S,P,B,3,3) tensor the per-scale `fmat` entries are slices of; when it requires grad its gradient
    is returned as ONE tensor instead of S*P slice gradients.
    cams / inv_Ks: poses (one (B,4,4) per pair) and inverse intrinsics (one (B,4,4) per scale) INSTEAD of fundamental
    matrices: F is built inside the call and the pose gradients come back from the same launches.
    axisangles / translations (+ inv_Ks): PoseNet's (B,1,1,3) outputs per pair INSTEAD of cams -- the call also runs
    transformation_from_parameters and returns the gradients w.r.t. the parameters."""
    library = library or _cabi.lib()
    slots, diff = [], []
    grad_on = torch.is_grad_enabled()
    poses = None
    if cams is not None:
        poses = ("cam", [c.detach() for c in cams], [k.detach() for k in inv_Ks])
        for p, c in enumerate(cams):
            if c.requires_grad and grad_on:
                slots.append((0, "cam", p))
                diff.append(c)
    elif axisangles is not None:
        poses = ("params", [a.detach() for a in axisangles], [t.detach() for t in translations], [k.detach() for k in inv_Ks])
        for kind, seq in (("aa", axisangles), ("tr", translations)):
            for p, t in enumerate(seq):
                if t.requires_grad and grad_on:
                    slots.append((0, kind, p))
                    diff.append(t)
    if fmat_all is not None and fmat_all.requires_grad and grad_on:
        slots.append((0, "fmat_all", 0))
        diff.append(fmat_all)
    for k, S in enumerate(scales):
        for kind in ("flow", "mob") + (() if (fmat_all is not None or cams is not None or axisangles is not None) else ("fmat",)):
            for p, t in enumerate(getattr(S, kind)):
                if t is not None and t.requires_grad and torch.is_grad_enabled():
                    if kind == "mob" and cfg.mask_mode == MASK_SHARED and p == 1:
                        continue
                    slots.append((k, kind, p))
                    diff.append(t)
    holder = {}
    if diff:
        total, terms = _FusedLossFn.apply(cfg, scales, slots, library, poses, *diff)
        maps = total.grad_fn.maps if total.grad_fn is not None and hasattr(total.grad_fn, "maps") else holder
    else:
        need = [{"flow": [False, False], "mob": [False, False], "fmat": [False, False]} for _ in scales]
        loss_out, _, maps, _ = run_fused(cfg, scales, need, library, poses=poses[1:] if poses and poses[0] == "cam" else None,
                                         pose_params=poses[1:] if poses and poses[0] == "params" else None)
        total, terms = loss_out[0], loss_out[1:5]
    return total, terms, maps
