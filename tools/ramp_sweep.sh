#!/bin/sh
# e2e (host-buffer) throughput for several chunk-ramp shapes (built on the GPU box)
for v in "-DP2V_RAMP_START_DIV=4" "-DP2V_RAMP_START_DIV=16" "-DP2V_RAMP_START_DIV=16 -DP2V_RAMP_NUM=5 -DP2V_RAMP_DEN=4" "-DP2V_RAMP_START_DIV=8 -DP2V_RAMP_NUM=5 -DP2V_RAMP_DEN=4" "-DP2V_RAMP_START_DIV=8"; do
  P2V_EXTRA_NVCC="$v" python plonky2-verifier_b200/build.py > /dev/null 2>&1 || { echo "build failed: $v"; continue; }
  echo "== $v"; python tools/e2e_probe.py 100000 2>&1 | tail -6 | head -3
done
python plonky2-verifier_b200/build.py > /dev/null 2>&1
