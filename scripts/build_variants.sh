#!/bin/bash
# Builds kernel-variant libraries HERE (CPU box; nvcc cross-compiles) into slam3d_b200/build/variants/ so that they travel with
# the gpurun snapshot.  usage: scripts/build_variants.sh name1 "<flags1>" name2 "<flags2>" ...
cd "$(dirname "$0")/.."
mkdir -p slam3d_b200/build/variants
while [ $# -ge 2 ]; do
  NAME=$1; FLAGS=$2; shift 2
  ( nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-O2 $FLAGS -shared -o slam3d_b200/build/variants/libs3d_$NAME.so \
      slam3d_b200/csrc/{voxel,grid,knn,gicp,ndt,map,api}.cu -lcudart 2>&1 | grep -E "error|warning: v" ; echo "built $NAME" ) &
done
wait
