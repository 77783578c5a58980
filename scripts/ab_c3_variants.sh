#!/bin/bash
# GPU box: scripts/c3_quick.py for the default library and for every variant under slam3d_b200/build/variants/ (or the names given)
mkdir -p gpurun_out
echo "== default"; timeout 120 python scripts/c3_quick.py 2>&1 | tail -3
NAMES=${1:-$(ls slam3d_b200/build/variants/ | sed 's/libs3d_//; s/\.so//')}
for N in $NAMES; do
  echo "== $N"; S3D_LIB_PATH=$PWD/slam3d_b200/build/variants/libs3d_$N.so timeout 120 python scripts/c3_quick.py 2>&1 | tail -3
done
