#!/bin/bash
# GPU box: what the driver runs at round end (GPU tests, smoke, both bench arms) + the secondary lines, outputs in gpurun_out/.
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}.err; tail -c 600 gpurun_out/bench_${TAG}_ref.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2>> gpurun_out/bench_${TAG}.err; tail -c 1500 gpurun_out/bench_${TAG}.json
timeout 300 python bench.py --workload c3 > gpurun_out/bench_${TAG}_c3.json 2>> gpurun_out/bench_${TAG}.err; tail -c 600 gpurun_out/bench_${TAG}_c3.json
timeout 300 python bench.py --workload c4 > gpurun_out/bench_${TAG}_c4.json 2>> gpurun_out/bench_${TAG}.err; tail -c 600 gpurun_out/bench_${TAG}_c4.json
timeout 300 python scripts/bench_ndt.py > gpurun_out/ndt_${TAG}.json 2>> gpurun_out/bench_${TAG}.err; tail -c 400 gpurun_out/ndt_${TAG}.json
timeout 200 python scripts/single_pair.py 2>&1 | tail -4
