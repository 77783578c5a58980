#!/bin/bash
# usage (on the GPU box, from the repo root): bash scripts/profile_round.sh r02
# Launch list of a short bench run + `ncu --set full` captures of the top kernels; the .ncu-rep files and the CSV land in
# gpurun_out/ and are summarised here (CPU) with scripts/profile_summary.py into profiles/<tag>_summary.md and with
# scripts/traffic_json.py into profiles/<tag>_traffic.json (the DRAM bytes per work unit bench.py scales `roofline.traffic` from).
TAG=${1:-r02}
B="python bench.py --steps 1 --warmup 1 --pairs 16 --distinct 4 --no-chain --no-cpu-baseline --no-extras"
export S3D_STREAMS_PER_DEVICE=3   # fewer, larger chunks per launch: the captured launches then hold 5-6 pairs each
export S3D_LOOP_MODE=3            # the per-pass kernels launched from the host: ncu does not see the kernel nodes inside a graph's WHILE body
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_${TAG}.csv $B > gpurun_out/b_launch_${TAG}.log 2>&1
for K in knn_cov_kernel gicp_search_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip 4 -c 1 -f -o gpurun_out/${K}_${TAG} $B > gpurun_out/b_${K}_${TAG}.log 2>&1
done
# the first search launch of a chunk: outer iteration 1, every point searches without a hint
ncu --set full --clock-control none --import-source on -k regex:gicp_search_kernel --launch-skip 0 -c 1 -f -o gpurun_out/gicp_search_kernel_first_${TAG} $B > gpurun_out/b_gicp_search_kernel_first_${TAG}.log 2>&1
# the same two kernels WITHOUT the cache flush between replays (--cache-control none): what L2 holds inside a running schedule
for K in knn_cov_kernel gicp_search_kernel; do
  ncu --set full --clock-control none --cache-control none -k regex:$K --launch-skip 4 -c 1 -f -o gpurun_out/${K}_warm_${TAG} $B > gpurun_out/b_${K}_warm_${TAG}.log 2>&1
done
# the persistent loop kernel of the single-call path (one pair) and the control kernel of the batch path
ncu --set full --clock-control none --import-source on -k regex:gicp_loop_kernel --launch-skip 2 -c 1 -f -o gpurun_out/gicp_loop_kernel_${TAG} env -u S3D_LOOP_MODE python scripts/single_pair.py > gpurun_out/b_gicp_loop_kernel_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gicp_ctrl_kernel --launch-skip 6 -c 1 -f -o gpurun_out/gicp_ctrl_kernel_${TAG} $B > gpurun_out/b_gicp_ctrl_kernel_${TAG}.log 2>&1
# config 3 (VoxelGrid on the 2M-point cloud, leaf 0.05 / 0.1 / 0.2 m, two calls each): time and DRAM bytes of every launch, with the
# cache flush between launches (cold) and without (warm: the 33.5 MB input stays in L2, as in a running pipeline)
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $M --clock-control none -c 60 --csv --log-file gpurun_out/c3_launches_${TAG}.csv python scripts/c3_voxel_only.py > /dev/null 2>&1
ncu --metrics $M --clock-control none --cache-control none -c 60 --csv --log-file gpurun_out/c3_launches_warm_${TAG}.csv python scripts/c3_voxel_only.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:sort_pass_kernel --launch-skip 4 -c 1 -f -o gpurun_out/sort_pass_kernel_${TAG} python scripts/c3_voxel_only.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:voxel_centroid_kernel --launch-skip 1 -c 1 -f -o gpurun_out/voxel_centroid_kernel_${TAG} python scripts/c3_voxel_only.py > /dev/null 2>&1
ls -la gpurun_out/*${TAG}*
