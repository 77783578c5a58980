#!/bin/bash
# GPU box: odometry-chain (device cache) and NDT throughput under different environment settings
mkdir -p gpurun_out
i=0
for E in "$@"; do
  i=$((i+1))
  env $(echo $E | tr ',' ' ') timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ab_chain_$i.json 2> gpurun_out/ab_chain_$i.err
  env $(echo $E | tr ',' ' ') timeout 300 python scripts/bench_ndt.py > gpurun_out/ab_chain_ndt_$i.json 2>> gpurun_out/ab_chain_$i.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/ab_chain_$i.json").read().strip().splitlines()[-1])
n = json.loads(open("gpurun_out/ab_chain_ndt_$i.json").read().strip().splitlines()[-1])
c = d["config"]["odometry_chain_device_cache"]
print("$E value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "chain", round(c["value"], 1), "chain ms", round(c["ms_per_step"], 2), "ndt", round(n["ndt_registrations_per_s_e2e"], 1))
PY
done
