"""Import the upstream reference IN PLACE (build container only) -- TEST INFRASTRUCTURE.

``/root/reference`` exists only in the build container, never on the GPU box:
nothing in ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may call this.  It is
used by ``oracle/make_golden.py`` (fixture generation) and by
``tests/test_oracle_vs_reference.py`` (which skips when the tree is absent).

Two shims are needed:
  * ``networks/__init__.py:4-5`` imports two modules that are not in the tree,
    so a stub ``networks`` package is pre-registered;
  * ``utils.py:3-11`` imports pypng / imageio / cityscapesscripts / detectron2,
    none of which is installed, so the three functions on the path are executed
    from their source slices instead of importing the module.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REF = os.environ.get("MDN_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "loss_functions.py"))


_cache = None


def load():
    """-> namespace with loss_functions, loss_utils, layers, and the utils.py slices."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF)
    sys.dont_write_bytecode = True  # the reference tree is read-only
    if REF not in sys.path:
        sys.path.insert(0, REF)
    if "networks" not in sys.modules:
        pkg = types.ModuleType("networks")
        pkg.__path__ = [os.path.join(REF, "networks")]
        sys.modules["networks"] = pkg
    layers = importlib.import_module("networks.layers")
    loss_utils = importlib.import_module("loss_utils")
    loss_functions = importlib.import_module("loss_functions")

    import numpy as np
    import torch
    from torch import nn
    src = open(os.path.join(REF, "utils.py")).read().split("\n")
    ns = {"np": np, "torch": torch, "nn": nn}

    def grab(first_line_startswith):
        start = next(k for k, l in enumerate(src) if l.startswith(first_line_startswith))
        end = start + 1
        while end < len(src) and (src[end].startswith((" ", "\t")) or src[end].strip() == ""):
            end += 1
        exec("\n".join(src[start:end]), ns)

    grab("def binary_image")
    grab("class FlowWarp")
    grab("def gauss_distance_weight")

    _cache = types.SimpleNamespace(
        loss_functions=loss_functions, loss_utils=loss_utils, layers=layers,
        binary_image=ns["binary_image"], FlowWarp=ns["FlowWarp"], gauss_distance_weight=ns["gauss_distance_weight"])
    return _cache


def load_networks():
    """Random-init FlowNet_v1 / PoseNet_v3 / MobileDecoder (trainer.py:139-142) for fixture inputs."""
    load()
    return (importlib.import_module("networks.flow_net_v1"), importlib.import_module("networks.pose_net_v3"),
            importlib.import_module("networks.mobile_decoder"))


def force_cpu_loss_module():
    """Make the reference LossModule default to cuda=False (loss_functions.py:12,18,172 default cuda=True).

    ``Loss.forward`` builds ``LossModule(self.opt, ssim=..., padding_mode=...)`` by its module-global name and
    the class calls ``super(LossModule, self)``, so the name cannot be rebound; the default is changed instead.
    """
    ref = load()
    ref.loss_functions.LossModule.__init__.__defaults__ = (None, None, "zeros", False)
    return ref.loss_functions.LossModule
