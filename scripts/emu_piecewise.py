"""Scratch check (build container, no GPU): runs the non-photometric tests of tests/test_gpu_piecewise.py on host tensors
under the SIMT emulator with shrunken shapes, so that their LOGIC is known to be right before a GPU round trip.
Test infrastructure only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from unittest import mock

import emu_harness
import test_gpu_piecewise as T
from mdn_sfm_b200 import loss_functions, loss_utils

T.DEV, T.SMALL = "cpu", True
cpu_arith = lambda cls: (lambda *a, **k: cls(*a, **dict(k, arith="cpu")))
with emu_harness.emulated():
    with mock.patch.object(loss_functions, "LossModule", cpu_arith(loss_functions.LossModule)), \
         mock.patch.object(loss_functions, "Loss", cpu_arith(loss_functions.Loss)):
        for mode in T.MODES:
            for form in ("list", "bare"):
                T.test_epipolar_loss_called_directly(mode, form)
                print("epipolar_loss direct", mode, form, "ok", flush=True)
        T.test_bare_instances_broadcast_over_a_batch()
        print("bare broadcast ok", flush=True)
        for mode in ("DC", "SN", "TG", "DS"):
            T.test_loss_module_forward_and_accumulators(mode)
            print("LossModule.forward", mode, "ok", flush=True)
        T.test_get_epipolar_new_on_point_sets()
        print("points ok", flush=True)
        T.test_compute_quantiles_product_function()
        print("quantiles ok", flush=True)
        T.test_single_source_frame_with_disable_min()
        print("single-frame disable_min ok", flush=True)
