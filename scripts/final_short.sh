#!/bin/bash
# GPU box: GPU tests, smoke and the bench line (the short form of final_round.sh)
TAG=${1:-r01i}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 1200 gpurun_out/bench_${TAG}.json
