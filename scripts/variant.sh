#!/bin/bash
# usage: scripts/variant.sh "<extra nvcc -D flags>" [bench args...]   -> builds a variant library in /tmp and prints the stage times
FLAGS="$1"; shift
OUT=/tmp/libs3d_variant_$$.so
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-O2 $FLAGS -shared -o $OUT slam3d_b200/csrc/voxel.cu slam3d_b200/csrc/grid.cu slam3d_b200/csrc/knn.cu slam3d_b200/csrc/gicp.cu slam3d_b200/csrc/ndt.cu slam3d_b200/csrc/map.cu slam3d_b200/csrc/api.cu -lcudart 2>&1 | grep -E "error" 
S3D_LIB_PATH=$OUT python bench.py --steps 3 --warmup 3 --pairs 32 --distinct 4 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$FLAGS', round(d['value']), {k:round(v,2) for k,v in d['roofline']['stage_ms_per_step'].items()})"
rm -f $OUT
