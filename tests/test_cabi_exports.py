"""The C-ABI library builds for sm_100a, loads without a GPU, and exports exactly what include/mdn_loss.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "mdn_loss.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mdn_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    from mdn_sfm_b200 import _cabi
    assert declared_symbols() == sorted(_cabi.EXPORTS)


def test_library_loads_and_exports_every_symbol():
    from mdn_sfm_b200 import _cabi, build
    path = build.build()
    dll = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(dll, name), name
    lib = _cabi.Library(path)
    assert lib.cdll.mdn_version() == _cabi.ABI_VERSION


def test_struct_layout_matches_header():
    """sizeof(MdnLossDesc) computed by a C compiler equals the ctypes mirror."""
    import subprocess
    import tempfile
    from mdn_sfm_b200 import _cabi
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "sz.c")
        open(c, "w").write('#include <stdio.h>\n#include <stddef.h>\n#include "mdn_loss.h"\nint main(){printf("%zu %zu %zu %zu",'
                           'sizeof(MdnScale),sizeof(MdnLossDesc),offsetof(MdnScale,g_flow),offsetof(MdnLossDesc,scale));return 0;}')
        exe = os.path.join(d, "sz")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        a, b, c_, d_ = map(int, subprocess.check_output([exe]).split())
    assert a == ctypes.sizeof(_cabi.MdnScale) and b == ctypes.sizeof(_cabi.MdnLossDesc)
    assert c_ == _cabi.MdnScale.g_flow.offset and d_ == _cabi.MdnLossDesc.scale.offset


def test_integration_stub_matches_the_header():
    """The ctypes structures printed in INTEGRATION.md (what a maintainer would copy) are EXECUTED and every field's offset
    and both sizes are held to the C compiler's view of include/mdn_loss.h; the block must also be what
    scripts/gen_integration_stub.py generates from the product's own binding."""
    import subprocess
    import sys
    import tempfile
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = doc[doc.index("<!-- stub:begin -->"):doc.index("<!-- stub:end -->")]
    code = block[block.index("```python") + len("```python"):block.rindex("```")]
    ns = {}
    exec(code, ns)
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import gen_integration_stub
    assert gen_integration_stub.stub().strip() in doc, "INTEGRATION.md stub is stale: run scripts/gen_integration_stub.py"
    items = []
    for cls in ("MdnScale", "MdnLossDesc"):
        items.append(("sizeof(%s)" % cls, ctypes.sizeof(ns[cls])))
        for name, _ in ns[cls]._fields_:
            items.append(("offsetof(%s,%s)" % (cls, name), getattr(ns[cls], name).offset))
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "off.c")
        body = "".join('printf("%%zu\\n", (size_t)%s);' % expr for expr, _ in items)
        open(c, "w").write('#include <stdio.h>\n#include <stddef.h>\n#include "mdn_loss.h"\nint main(){%s return 0;}' % body)
        exe = os.path.join(d, "off")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        got = [int(v) for v in subprocess.check_output([exe]).split()]
    for (expr, mine), theirs in zip(items, got):
        assert mine == theirs, (expr, mine, theirs)


def test_product_refuses_cpu_tensors():
    import torch
    from mdn_sfm_b200 import synthetic
    from mdn_sfm_b200.loss_functions import Loss
    opt = synthetic.default_opt(1, 16, 32)
    inputs, flows, mobiles, cams, inst = synthetic.make_batch(1, 16, 32, scales=(0,))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        Loss(opt, mode="T")(inputs, [-1, 1], flows, mobiles, inst, [0], cams)


def test_error_codes():
    from mdn_sfm_b200 import _cabi, build
    lib = _cabi.Library(build.build())
    call = _cabi.FusedCall(batch=0, n_pairs=2, post=0, mask_mode=0, flags=1)
    assert lib.cdll.mdn_loss_workspace_bytes(ctypes.byref(call.desc)) == 0
    assert b"out of range" in lib.cdll.mdn_last_error_string()
    with pytest.raises(RuntimeError, match="NULL"):
        lib.call("mdn_ssim_fwd", None, None, None, 1, 4, 4, None)
    with pytest.raises(RuntimeError, match="shape"):
        lib.call("mdn_flow_warp_bwd", 16, 16, 16, 16, 1, 3, 1, 8, 0, None)


def test_error_codes_of_the_data_side_entry_points():
    """Argument checks of the instance-mask / pyramid entry points and of the pose inputs of mdn_loss_fused return status
    codes before any launch (no GPU needed)."""
    import ctypes as C
    from mdn_sfm_b200 import _cabi, build
    lib = _cabi.Library(build.build())
    one = (C.c_int32 * 1)(8)
    with pytest.raises(RuntimeError, match="NULL"):
        lib.call("mdn_instance_mask_union", None, None, None, 1, 16, None)
    with pytest.raises(RuntimeError, match="out of range"):
        lib.call("mdn_instance_mask_resize", 16, 0, 8, 8, (C.c_void_p * 1)(16), one, one, 1, 16, 1 << 20, None)
    with pytest.raises(RuntimeError, match="workspace too small"):
        lib.call("mdn_image_pyramid", 16, 3, 8, 8, (C.c_void_p * 1)(16), one, one, 1, 16, 0, None)
    with pytest.raises(RuntimeError, match="multiple of 3"):
        lib.call("mdn_image_pyramid_packed", 16, 4, 8, 8, (C.c_void_p * 1)(16), one, one, 1, 16, 1 << 20, None)
    assert lib.cdll.mdn_instance_mask_resize_workspace_bytes(2, 8, 8, one, one, 1) > 0
    assert lib.cdll.mdn_instance_mask_resize_workspace_bytes(0, 8, 8, one, one, 1) == 0
    # poses: every pair's pose or none, and the inverse intrinsics of every scale
    call = _cabi.FusedCall(batch=1, n_pairs=2, post=1, mask_mode=2, flags=1)
    call.desc.n_scales = 1
    call.desc.scale[0].height, call.desc.scale[0].width = 8, 8
    call.desc.cam[0] = 16
    assert lib.cdll.mdn_loss_workspace_bytes(C.byref(call.desc)) == 0
    assert b"every pair" in lib.cdll.mdn_last_error_string()
    call.desc.cam[1] = 16
    assert lib.cdll.mdn_loss_workspace_bytes(C.byref(call.desc)) == 0
    assert b"inv_K" in lib.cdll.mdn_last_error_string()
