#!/bin/sh
# analysis only: K6a time of the what-if builds (P2V_WHATIF: wrong results on purpose) -> marginal cost of each component
for so in plonky2-verifier_b200/variants/libp2v_w*.so; do
  P2V_LIB_PATH=$PWD/$so timeout 300 python tools/perf_k6a.py ${1:-40000} 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%-16s k6a %.2f ms  challenges %.2f ms  step %.2f ms' % (d['lib'], d['kernel_ms']['fri_merkle'], d['kernel_ms']['challenges'], d['step_ms']))"
done
