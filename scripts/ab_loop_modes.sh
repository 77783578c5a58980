#!/bin/bash
# GPU box: the GICP loop in its two modes (S3D_LOOP_MODE 1 = one persistent launch, 2 = graph replay with non-waiting CTAs) and
# the linger setting of the graph mode, on the default bench workload.
mkdir -p gpurun_out
for CFG in "0 8" "1 8" "2 0" "2 8" "2 32"; do
  set -- $CFG
  S3D_LOOP_MODE=$1 S3D_LOOP_LINGER=$2 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ab_loop_$1_$2.json 2> gpurun_out/ab_loop_$1_$2.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/ab_loop_$1_$2.json").read().strip().splitlines()[-1])
print("mode $1 linger $2: value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 2), "launches", d["gpu_launches"],
      {k: round(v, 2) for k, v in d["roofline"]["stage_ms_per_step"].items()}, "chain", round(d["config"].get("odometry_chain_device_cache", {}).get("value", 0)))
PY
done
