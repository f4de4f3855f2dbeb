"""Out-of-bounds check of the kernel source on the CPU: the SIMT-emulated build (tests/emu) compiled with AddressSanitizer and
driven through the product's Python layer over ragged / tiny shapes, every mode, the maps-only launch, the instance-mask
and pyramid kernels.  compute-sanitizer is closed on the GPU pool; host tensors and the emulated shared memory are heap
blocks ASan guards.  Usage (see scripts/asan_emu.sh):
    LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python scripts/asan_emu_run.py /tmp/libmdn_loss_emu_asan.so
TEST INFRASTRUCTURE -- the product never loads the emulated build."""
import os
import sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import common
from unittest import mock
from mdn_sfm_b200 import _cabi, loss_utils, pyramid, synthetic
lib = _cabi.Library(sys.argv[1])
def check_tensor(t, dtype=None, what="tensor"):
    return t
with mock.patch.object(_cabi, "_lib", lib), mock.patch.object(_cabi, "check_tensor", check_tensor):
    for (B,H,W,scales,mode,photo) in [(1,23,45,(0,),"SN",True),(2,16,70,(0,1),"DC",True),(1,33,64,(0,),"TG",False),(1,5,3,(0,),"T",True)]:
        opt,batch=common.make(B,H,W,scales=scales,seed=5,flow_std=0.2)
        got=common.product_run(opt,batch,mode,photo,True,"cpu",pose_grad=True,arith="cuda")
        _ = got[0]["epipolars"][(-1,0)].sum()
        print(mode, float(got[1]["loss"]))
    img=torch.rand(1,3,20,36)
    print([t.shape for t in pyramid.image_pyramid(img,[(10,18),(7,5)],lib)])
print("done")
