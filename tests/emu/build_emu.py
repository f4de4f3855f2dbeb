"""Builds the SIMT-emulated copy of the kernels (TEST INFRASTRUCTURE -- see cuda_emu.h).

g++ compiles the unmodified mdn_sfm_b200/csrc/mdn_loss.cu against tests/emu/cuda_emu.h into
tests/emu/_build/libmdn_loss_emu.so, which only the not-gpu tests load.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(ROOT, "mdn_sfm_b200", "csrc", "mdn_loss.cu")
DEPS = [SRC, os.path.join(ROOT, "mdn_sfm_b200", "csrc", "mdn_common.cuh"), os.path.join(os.path.join(ROOT, "mdn_sfm_b200"), "csrc", "mdn_fused.cuh"), os.path.join(ROOT, "mdn_sfm_b200", "csrc", "mdn_resize.cuh"), os.path.join(ROOT, "include", "mdn_loss.h"),
        os.path.join(HERE, "cuda_emu.h")]
OUT = os.path.join(HERE, "_build", "libmdn_loss_emu.so")


def build(force=False):
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-x", "c++", "-std=c++17", "-O2", "-g", "-ffp-contract=off", "-fno-fast-math", "-DMDN_EMU=1",
           "-fvisibility=default", "-shared", "-fPIC", "-Wno-attributes", "-I", HERE, "-I", os.path.join(ROOT, "include"),
           "-o", OUT, SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("emu build failed")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
