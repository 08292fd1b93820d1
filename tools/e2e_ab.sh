#!/bin/sh
# tuning aid: device-resident and end-to-end throughput of bench.py for a list of prebuilt variants x P2V_SPLIT settings
for so in "$@"; do
  for sp in 0 1; do
    P2V_SPLIT=$sp P2V_LIB_PATH=$PWD/plonky2-verifier_b200/variants/libp2v_$so.so python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$so split=$sp value %.0f e2e %.0f frac %.4f k6a %.2f' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['kernel_ms']['fri_merkle']))"
  done
done
