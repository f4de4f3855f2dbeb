#!/bin/bash
# builds a tuning variant of libmdn_loss.so: scripts/build_variant.sh <name> [-D... nvcc flags]   -> scratch/variants/<name>.so
set -e
name=$1; shift
mkdir -p scratch/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
  -Xptxas -v --expt-relaxed-constexpr "$@" -I include -o scratch/variants/$name.so mdn_sfm_b200/csrc/mdn_loss.cu 2>&1 \
  | grep -A2 "fused_tile_kernelILb1ELb0" | grep -E "spill|registers" | sed "s/^/$name: /"
