#!/bin/bash
# ncu --set full on one conv shape. usage: gpu_ncu_conv.sh "B D cin cout" impl tag
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
shape=${1:-"32 32 32 64"}; impl=${2:-halo}; tag=${3:-c2}
ICSG3D_CONV_IMPL=$impl python tools/conv_case.py $shape 5
ICSG3D_CONV_IMPL=$impl timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3d_k3 -s 2 -c 1 \
   -f -o gpurun_out/ncu_${tag}_$impl python tools/conv_case.py $shape 2 > gpurun_out/ncu_${tag}_$impl.log 2>&1
tail -2 gpurun_out/ncu_${tag}_$impl.log
