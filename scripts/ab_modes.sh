#!/bin/bash
# A/B of the search-kernel variants on the GPU box.  usage: bash scripts/ab_modes.sh "K,N K,N ..."  (S3D_KNN_MODE,S3D_NN_MODE pairs)
mkdir -p gpurun_out
[ -n "$SKIP_TESTS" ] || timeout 900 python -m pytest tests/test_gpu_nn.py tests/test_gpu_gicp.py -x -q -m gpu 2>&1 | tail -3
MODES=${1:-0,0 1,0 0,1 1,1}
for M in $MODES; do
  K=${M%,*}; N=${M#*,}
  S3D_KNN_MODE=$K S3D_NN_MODE=$N timeout 300 python bench.py --steps 6 --warmup 3 --no-chain --no-cpu-baseline > gpurun_out/ab_mode_${K}_${N}.json 2> gpurun_out/ab_mode_${K}_${N}.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/ab_mode_${K}_${N}.json").read().strip().splitlines()[-1])
print("KNN_MODE=$K NN_MODE=$N value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 2), {k: round(v, 2) for k, v in d["roofline"]["stage_ms_per_step"].items()})
PY
done
