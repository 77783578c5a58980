#!/bin/bash
# GPU box: bench under different environment settings.  usage: bash scripts/ab_env.sh "NAME=VAL,NAME2=VAL2" "..." ...
mkdir -p gpurun_out
i=0
for E in "$@"; do
  i=$((i+1))
  env $(echo $E | tr ',' ' ') timeout 300 python bench.py --steps 8 --warmup 3 --no-chain --no-cpu-baseline > gpurun_out/ab_env_$i.json 2> gpurun_out/ab_env_$i.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/ab_env_$i.json").read().strip().splitlines()[-1])
print("$E value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 2), "e2e ms", round(d["e2e"]["ms_per_step"], 2))
PY
done
