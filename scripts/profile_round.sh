#!/bin/bash
# usage (on the GPU box, from the repo root): bash scripts/profile_round.sh r01h
# Launch list of a short bench run + one `ncu --set full` capture of the top kernels; the .ncu-rep files and the CSV land in
# gpurun_out/ and are summarised here (CPU) with scripts/profile_summary.py into profiles/<tag>_summary.md.
TAG=${1:-r01h}
B="python bench.py --steps 1 --warmup 1 --pairs 16 --distinct 4 --no-chain --no-cpu-baseline"
export S3D_STREAMS_PER_DEVICE=3   # fewer, larger chunks per launch: the captured launches then hold 5-6 pairs each
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_${TAG}.csv $B > gpurun_out/b_launch_${TAG}.log 2>&1
for K in knn_cov_kernel gicp_iter_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip 4 -c 1 -f -o gpurun_out/${K}_${TAG} $B > gpurun_out/b_${K}_${TAG}.log 2>&1
done
# the first gicp_iter_kernel launch of a chunk: outer iteration 1, every point searches without a hint
ncu --set full --clock-control none --import-source on -k regex:gicp_iter_kernel --launch-skip 0 -c 1 -f -o gpurun_out/gicp_iter_kernel_first_${TAG} $B > gpurun_out/b_gicp_iter_kernel_first_${TAG}.log 2>&1
if [ -z "$SKIP_NDT" ]; then
ncu --set full --clock-control none --import-source on -k regex:ndt_eval_kernel --launch-skip 6 -c 1 -f -o gpurun_out/ndt_eval_kernel_${TAG} python scripts/bench_ndt.py --pairs 16 --steps 1 > gpurun_out/b_ndt_eval_kernel_${TAG}.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_ndt_${TAG}.csv python scripts/bench_ndt.py --pairs 16 --steps 1 > gpurun_out/b_launch_ndt_${TAG}.log 2>&1
fi
ls -la gpurun_out/*${TAG}*
