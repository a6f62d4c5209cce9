#!/bin/bash
# Builds libdvo.so (sm_100a) in-tree.  -fmad=false: every float contraction in the bit-exact stages is explicit.
set -e
cd "$(dirname "$0")"
SRC=droplet_visual_odometry_b200/csrc
OUT=droplet_visual_odometry_b200/libdvo.so
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false \
     -Xcompiler -fPIC -shared -Iinclude -o $OUT $SRC/dvo_api.cu $SRC/orb_kernels.cu $SRC/pair_kernels.cu $SRC/nn_tensor.cu $SRC/peaks.cu "$@"
echo "built $OUT"
