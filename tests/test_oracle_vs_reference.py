"""Pins oracle/restate.py bit-exactly against the upstream reference imported in place.

Runs only where /root/reference exists (the build container); skipped on the GPU box.
"""
import copy

import pytest
import torch

from mdn_sfm_b200 import synthetic
from oracle import ref_loader, ref_modes, restate

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")

B, H, W = 2, 32, 64


def _leafify(flows, mobiles):
    flows = {k: v.clone().requires_grad_(True) for k, v in flows.items()}
    mobiles = {k: v.clone().requires_grad_(True) for k, v in mobiles.items()}
    return flows, mobiles


@pytest.mark.parametrize("mode,photo,ssim_on,disable_min", [
    ("DC", False, False, False), ("SN", True, True, False), ("T", True, True, False), ("TG", True, False, True),
    ("DS", False, False, False), ("DC", True, True, True)])
def test_loss_forward_and_grads_bit_exact(mode, photo, ssim_on, disable_min):
    opt = synthetic.default_opt(B, H, W, disable_min=disable_min)
    inputs, flows, mobiles, cams, inst = synthetic.make_batch(B, H, W, seed=7, flow_std=0.05)
    weights = restate.gauss_distance_weight(4, H, W) if mode == "TG" else None
    f1, m1 = _leafify(flows, mobiles)
    f2, m2 = _leafify(flows, mobiles)
    o_ref, l_ref = ref_modes.reference_loss_forward(opt, inputs, [-1, 1], f1, m1, inst, [0, 1, 2, 3], cams, mode=mode,
                                                    weights=weights, photometric=photo, ssim_on=ssim_on)
    o_or, l_or = restate.loss_forward(opt, inputs, [-1, 1], f2, m2, inst, [0, 1, 2, 3], cams, mode=mode,
                                      weights=weights, photometric=photo, ssim_on=ssim_on)
    for k in ("loss", "epip", "smooth", "consis"):
        assert torch.equal(l_ref[k], l_or[k]), k
    l_ref["loss"].backward()
    l_or["loss"].backward()
    for k in f1:
        assert torch.equal(f1[k].grad, f2[k].grad), k
    for k in m1:
        assert torch.equal(m1[k].grad, m2[k].grad), k
    for name in ("epipolars", "epipolar_ori", "flows") + (("warps", "diffs", "valids") if photo else ()):
        for key in o_ref[name]:
            assert torch.equal(o_ref[name][key], o_or[name][key]), (name, key)
    for s in o_ref["min_mobiles"]:
        assert torch.equal(o_ref["min_mobiles"][s], o_or["min_mobiles"][s])


def test_free_functions_bit_exact():
    ref = ref_loader.load()
    g = torch.Generator().manual_seed(3)
    aa = torch.randn(3, 1, 1, 3, generator=g) * 0.3
    tt = torch.randn(3, 1, 1, 3, generator=g)
    for inv in (False, True):
        assert torch.equal(ref.layers.transformation_from_parameters(aa, tt, inv),
                           restate.transformation_from_parameters(aa, tt, inv))
    assert torch.equal(ref.layers.get_scale_factor(2, 5, 7).contiguous(), restate.get_scale_factor(2, 5, 7).contiguous())
    assert torch.equal(ref.loss_utils.create_coords(2, 5, 7), restate.create_coords(2, 5, 7))
    x = torch.rand(2, 3, 9, 11, generator=g)
    y = torch.rand(2, 3, 9, 11, generator=g)
    assert torch.equal(ref.layers.SSIM()(x, y), restate.ssim(x, y))
    assert torch.equal(ref.binary_image(x, 0.4), restate.binary_image(x, 0.4))
    fl = torch.randn(2, 2, 9, 11, generator=g) * 3
    a, b = ref.FlowWarp(2, 9, 11)(fl), restate.flow_warp_grid(fl)
    for u, v in zip(a, b):
        assert torch.equal(u, v)
    for wr, wo in zip(ref.gauss_distance_weight(3, 32, 64), restate.gauss_distance_weight(3, 32, 64)):
        assert torch.equal(wr, wo)
    m = torch.rand(2, 1, 9, 11, generator=g)
    assert torch.equal(ref.loss_utils.smooth_loss(x, m), restate.smooth_loss(x, m))
    assert torch.equal(ref.loss_utils.derivable_consistency_loss(m, 1 - m), restate.derivable_consistency_loss(m, 1 - m))


def test_bare_instances_branch():
    ref = ref_loader.load()
    g = torch.Generator().manual_seed(5)
    inst = synthetic.make_instances(1, g)[0]["instances"]
    assert torch.equal(ref.loss_utils.get_batch_instance_mask(inst), restate.get_batch_instance_mask(inst))
