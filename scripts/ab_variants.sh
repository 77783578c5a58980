#!/bin/bash
# GPU box: bench every variant library under slam3d_b200/build/variants/ (or the names given) and print the stage times.
mkdir -p gpurun_out
NAMES=${1:-$(ls slam3d_b200/build/variants/ | sed 's/libs3d_//; s/\.so//')}
for N in $NAMES; do
  S3D_LIB_PATH=$PWD/slam3d_b200/build/variants/libs3d_$N.so timeout 300 python bench.py --steps 6 --warmup 3 --no-chain --no-cpu-baseline > gpurun_out/ab_var_$N.json 2> gpurun_out/ab_var_$N.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/ab_var_$N.json").read().strip().splitlines()[-1])
print("$N value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 2), {k: round(v, 2) for k, v in d["roofline"]["stage_ms_per_step"].items()})
PY
done
