#!/bin/sh
# build + bench a list of compile-time variants on the GPU box (needs nvcc there: the image has it)
for v in "$@"; do
  P2V_EXTRA_NVCC="$v" python plonky2-verifier_b200/build.py > /dev/null 2>&1 || { echo "build failed: $v"; continue; }
  printf "%s : " "$v"; tools/quick_bench.sh 50000
done
