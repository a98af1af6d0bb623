#!/bin/bash
# every conv shape of the VAE+DFC step (batch 32) under the three kernels: default dispatch, forced halo, forced per-tap
cd "$(dirname "$0")/.."
for shape in "32 32 16 32" "32 32 32 16" "32 32 32 64" "32 32 64 32" "32 16 64 64" "32 16 64 128" "32 16 128 64" "32 8 128 128" \
  "32 8 128 256" "32 8 256 128" "32 4 256 512" "32 4 512 256" "32 4 512 512" "32 32 48 16" "32 16 48 32" "32 8 96 64" "32 4 192 128" \
  "32 2 384 16" "32 16 32 16" "32 8 64 32" "32 4 128 64" "32 2 16 128" "32 4 16 128" "32 8 128 64" "32 16 64 32" "32 32 16 16" \
  "32 4 128 16" "32 8 64 128" "32 16 32 64"; do
  for impl in auto halo v1; do ICSG3D_CONV_IMPL=$impl python tools/conv_case.py $shape 20 2>&1 | tail -1; done
done
