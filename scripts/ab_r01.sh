#!/bin/bash
# GPU box: round-1 library (slam3d_b200/build/variants/libs3d_r01.so, built from commit 62c3a9d) against the tree, same workload.
mkdir -p gpurun_out
for S in 3 6; do for M in 2 3; do S3D_STREAMS_PER_DEVICE=$S S3D_LOOP_MODE=$M python scripts/loop_only.py; done; done
for S in 6; do for M in 3; do S3D_BLOCKING_SYNC=0 S3D_STREAMS_PER_DEVICE=$S S3D_LOOP_MODE=$M python scripts/loop_only.py; done; done
