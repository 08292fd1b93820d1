#!/bin/sh
for c in 1536 2048 3072 4096 6144 8160; do
  echo "== chunk $c"; python tools/e2e_probe.py 100000 $c 2>&1 | tail -7 | head -4
done
P2V_EXTRA_NVCC="-DP2V_RAMP_NUM=17 -DP2V_RAMP_DEN=16" python plonky2-verifier_b200/build.py > /dev/null 2>&1
echo "== ramp 17/16"; python tools/e2e_probe.py 100000 2>&1 | tail -7 | head -4
python plonky2-verifier_b200/build.py > /dev/null 2>&1
