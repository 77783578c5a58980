#!/bin/bash
# GPU box: the default bench workload under scheduling knobs (and the round-1 library for reference)
mkdir -p gpurun_out
run() {  # name, env...
  NAME=$1; shift
  env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-chain --no-extras > gpurun_out/abf_$NAME.json 2> gpurun_out/abf_$NAME.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/abf_$NAME.json").read().strip().splitlines()[-1])
    print("$NAME: value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "pageable", round(d["e2e_pageable"]["value"], 1), "ms/step", round(d["ms_per_step"], 2),
          "launches", d["gpu_launches"], {k: round(v, 2) for k, v in d["roofline"]["stage_ms_per_step"].items()})
except Exception as e:
    print("$NAME: failed", e)
PY
}
for S in ${STREAMS:-3 4 5 6}; do run streams$S S3D_STREAMS_PER_DEVICE=$S; done
run default A=1
