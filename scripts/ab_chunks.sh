#!/bin/bash
# e2e / value against the chunking knobs.  usage: bash scripts/ab_chunks.sh "P,W P,W ..."  (S3D_MAX_PAIRS_PER_LAUNCH,S3D_STREAMS_PER_DEVICE)
mkdir -p gpurun_out
COMBOS=${1:-32,3 16,3 11,3 8,3 16,4 11,4 8,4 11,6 6,6}
for M in $COMBOS; do
  P=${M%,*}; W=${M#*,}
  S3D_KNN_MODE=0 S3D_NN_MODE=0 S3D_MAX_PAIRS_PER_LAUNCH=$P S3D_STREAMS_PER_DEVICE=$W timeout 300 python bench.py --steps 8 --warmup 3 --no-chain --no-cpu-baseline > gpurun_out/ab_chunk_${P}_${W}.json 2> gpurun_out/ab_chunk_${P}_${W}.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/ab_chunk_${P}_${W}.json").read().strip().splitlines()[-1])
print("PAIRS=$P STREAMS=$W value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 2), "e2e ms", round(d["e2e"]["ms_per_step"], 2))
PY
done
