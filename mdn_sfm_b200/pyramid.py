"""Image-pyramid producer on the device: the step in front of the loss on the data side.

The reference's dataset resizes every frame to the four pyramid levels on the CPU and uploads all of them
(mono_dataset.py:106-125, trainer.py:226-227).  ``image_pyramid`` makes the lower levels from the full-resolution frame
on the GPU -- torchvision ``Resize`` semantics (bilinear + antialias on an fp32 tensor), ATen's separable filter replayed by
``mdn_image_pyramid`` -- so only the full-resolution frames cross PCIe (a quarter fewer image bytes per step).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi


def image_pyramid(img, sizes, library=None):
    """img (B,C,H,W) fp32 CUDA tensor -> [Resize(size)(img) for size in sizes] in one call (three launches).

    Sizes equal to (H, W) return `img` itself, like torchvision does."""
    library = library or _cabi.lib()
    img = _cabi.check_tensor(img, what="img").contiguous()
    B, Cn, H, W = img.shape
    outs = [img if (int(h), int(w)) == (H, W) else torch.empty((B, Cn, int(h), int(w)), dtype=torch.float32, device=img.device)
            for h, w in sizes]
    todo = [o for o in outs if o is not img]
    for k0 in range(0, len(todo), _cabi.MAX_SCALES):
        chunk = todo[k0:k0 + _cabi.MAX_SCALES]
        oh = (C.c_int32 * len(chunk))(*[o.shape[2] for o in chunk])
        ow = (C.c_int32 * len(chunk))(*[o.shape[3] for o in chunk])
        nbytes = library.cdll.mdn_instance_mask_resize_workspace_bytes(B * Cn, H, W, oh, ow, len(chunk))
        ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=img.device)
        library.call("mdn_image_pyramid", img.data_ptr(), B * Cn, H, W, _cabi.ptr_array(chunk), oh, ow, len(chunk),
                     ws.data_ptr(), nbytes, _cabi.stream_ptr(img))
    return outs


def add_pyramid_levels(inputs, frame_ids, scales, library=None):
    """Fills inputs[("color", i, s)] for s >= 1 from inputs[("color", i, 0)] (what the dataset's per-scale Resize produced)."""
    for i in frame_ids:
        full = inputs[("color", i, 0)]
        H, W = full.shape[-2:]
        lower = [s for s in scales if s != 0]
        for s, t in zip(lower, image_pyramid(full, [(H // 2 ** s, W // 2 ** s) for s in lower], library)):
            inputs[("color", i, s)] = t
    return inputs
