#!/bin/bash
# time the VAE+DFC layer shapes (B=32) with the conv implementations. usage: gpu_conv_sweep.sh [impls]
cd "$(dirname "$0")/.."
impls=${1:-"v1 halo"}
for shape in "32 32 16 32" "32 32 32 16" "32 32 32 64" "32 32 64 32" "32 32 16 16" "32 16 64 64" "32 16 64 128" "32 16 128 64" "32 16 64 32" "32 8 128 128" "32 8 128 256" "32 8 256 128" "32 8 128 64"; do
  for impl in $impls; do
    ICSG3D_CONV_IMPL=$impl python tools/conv_case.py $shape 10 2>&1 | tail -1
  done
done
