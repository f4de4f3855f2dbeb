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


def frames_from_u8(frames_u8, mean=(0.45, 0.45, 0.45), std=(0.225, 0.225, 0.225), library=None):
    """(B,H,W,3) uint8 CUDA frames, as the loader holds them -> (B,3,H,W) fp32 normalised frames: the dataset's
    ArrayToTensor + Normalize (custom_transforms.py:72-80,103-112; mean / std of mono_dataset.py:51-52) on the device,
    bit-identical to the CPU ops.  Frames then cross PCIe as bytes."""
    library = library or _cabi.lib()
    src = _cabi.check_tensor(frames_u8, dtype=torch.uint8, what="frames_u8").contiguous()
    if src.dim() != 4 or src.shape[-1] != 3:
        raise ValueError("frames_u8 must be (B,H,W,3)")
    B, H, W, _ = src.shape
    out = torch.empty((B, 3, H, W), dtype=torch.float32, device=src.device)
    m, s = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
    library.call("mdn_normalize_u8", src.data_ptr(), out.data_ptr(), B, H, W, m, s, _cabi.stream_ptr(src), dev=src)
    return out


def image_pyramid(img, sizes, library=None, packed=False):
    """img (B,C,H,W) fp32 CUDA tensor -> [Resize(size)(img) for size in sizes] in one call (three launches).

    Sizes equal to (H, W) return `img` itself, like torchvision does.  packed=True (C == 3): every level comes as
    (B,h,w,4), one (r, g, b, 0) float4 per pixel -- the layout the warp gather of the loss reads, handed to `Loss.forward`
    as inputs[("color_packed", i, s)] so that it launches no repack kernel; (H, W) itself is then a pure repack."""
    library = library or _cabi.lib()
    img = _cabi.check_tensor(img, what="img").contiguous()
    B, Cn, H, W = img.shape
    if packed:
        if Cn != 3:
            raise ValueError("packed pyramids need 3 channels")
        outs = [torch.empty((B, int(h), int(w), 4), dtype=torch.float32, device=img.device) for h, w in sizes]
        for k0 in range(0, len(outs), _cabi.MAX_SCALES):
            chunk = outs[k0:k0 + _cabi.MAX_SCALES]
            oh = (C.c_int32 * len(chunk))(*[o.shape[1] for o in chunk])
            ow = (C.c_int32 * len(chunk))(*[o.shape[2] for o in chunk])
            nbytes = library.cdll.mdn_instance_mask_resize_workspace_bytes(B * 3, H, W, oh, ow, len(chunk))
            ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=img.device)
            library.call("mdn_image_pyramid_packed", img.data_ptr(), B * 3, H, W, _cabi.ptr_array(chunk), oh, ow, len(chunk),
                         ws.data_ptr(), nbytes, _cabi.stream_ptr(img), dev=img)
        return outs
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
                     ws.data_ptr(), nbytes, _cabi.stream_ptr(img), dev=img)
    return outs


def add_pyramid_levels(inputs, frame_ids, scales, library=None, packed_sources=False):
    """Fills inputs[("color", i, s)] for s >= 1 from inputs[("color", i, 0)] (what the dataset's per-scale Resize produced).

    packed_sources=True: the source frames (i != 0) get inputs[("color_packed", i, s)] for EVERY scale instead -- the loss
    then runs without its repack kernel (the NCHW levels of the source frames are not needed by it)."""
    for i in frame_ids:
        full = inputs[("color", i, 0)]
        H, W = full.shape[-2:]
        if packed_sources and i != 0:
            for s, t in zip(scales, image_pyramid(full, [(H // 2 ** s, W // 2 ** s) for s in scales], library, packed=True)):
                inputs[("color_packed", i, s)] = t
            continue
        lower = [s for s in scales if s != 0]
        for s, t in zip(lower, image_pyramid(full, [(H // 2 ** s, W // 2 ** s) for s in lower], library)):
            inputs[("color", i, s)] = t
    return inputs
