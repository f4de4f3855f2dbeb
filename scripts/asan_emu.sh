#!/bin/bash
# Builds the SIMT-emulated kernels with AddressSanitizer and runs scripts/asan_emu_run.py under it (CPU only).
set -e
cd "$(dirname "$0")/.."
OUT=${TMPDIR:-/tmp}/libmdn_loss_emu_asan.so
g++ -x c++ -std=c++17 -O1 -g -fsanitize=address -fno-omit-frame-pointer -ffp-contract=off -fno-fast-math -DMDN_EMU=1 \
  -fvisibility=default -shared -fPIC -Wno-attributes -I tests/emu -I include -o "$OUT" mdn_sfm_b200/csrc/mdn_loss.cu
ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0:verify_asan_link_order=0 \
  LD_PRELOAD=$(gcc -print-file-name=libasan.so) python scripts/asan_emu_run.py "$OUT"
