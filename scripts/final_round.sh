#!/bin/bash
# GPU box: what the driver runs at round end (GPU tests, smoke, both bench arms) + the secondary configs, outputs in gpurun_out/.
TAG=${1:-r01h}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 1500 gpurun_out/bench_${TAG}.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2>> gpurun_out/bench_${TAG}.err; tail -c 600 gpurun_out/bench_${TAG}_ref.json
timeout 300 python scripts/bench_c4.py > gpurun_out/c4_${TAG}.json 2>> gpurun_out/bench_${TAG}.err; tail -c 600 gpurun_out/c4_${TAG}.json
timeout 300 python scripts/bench_ndt.py > gpurun_out/ndt_${TAG}.json 2>> gpurun_out/bench_${TAG}.err; tail -c 400 gpurun_out/ndt_${TAG}.json
