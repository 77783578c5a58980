#!/bin/bash
# A/B of the two kNN kernels on the GPU box: parity tests with each, then the bench stage times with each.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nn.py tests/test_gpu_gicp.py -x -q -m gpu 2>&1 | tail -5
for T in 0 1; do
  S3D_KNN_TILE=$T timeout 300 python bench.py --steps 6 --warmup 3 --no-chain --no-cpu-baseline > gpurun_out/ab_knn_tile_$T.json 2> gpurun_out/ab_knn_tile_$T.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/ab_knn_tile_$T.json").read().strip().splitlines()[-1])
print("TILE=$T value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 2), {k: round(v, 2) for k, v in d["roofline"]["stage_ms_per_step"].items()})
PY
done
