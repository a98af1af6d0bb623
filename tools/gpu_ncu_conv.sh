#!/bin/bash
# ncu --set full on the conv kernels (c2 fprop shape), both implementations
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for impl in v1 halo; do
  ICSG3D_CONV_IMPL=$impl python tools/conv_case.py 32 32 32 64 5
  ICSG3D_CONV_IMPL=$impl timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3d_k3 -s 2 -c 1 \
     -f -o gpurun_out/ncu_c2_$impl python tools/conv_case.py 32 32 32 64 2 > gpurun_out/ncu_c2_$impl.log 2>&1
  tail -3 gpurun_out/ncu_c2_$impl.log
done
ls -la gpurun_out/
