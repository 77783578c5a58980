#!/bin/bash
# ncu capture of the tile kNN kernel + work statistics (stats build shipped in slam3d_b200/build/)
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --pairs 16 --distinct 4 --no-chain --no-cpu-baseline"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:knn_cov --launch-skip 4 -c 1 -f -o gpurun_out/knn_cov_tile_${1:-r01h} $B > gpurun_out/b_knn_cov_tile_${1:-r01h}.log 2>&1
for T in 0 1; do echo "TILE=$T"; S3D_KNN_TILE=$T timeout 200 python scripts/knn_stats.py 2>&1 | tail -2; done
