#!/bin/sh
# tuning aid: bench.py (device-resident value + end-to-end) for launch-priority modes x tail taper of the host schedule
for pr in ${PRIOS:-0 1 2}; do
  for tl in ${TAILS:-0 1056}; do
    P2V_PRIO=$pr P2V_TAIL=$tl python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('prio=$pr tail=$tl value %.0f e2e %.0f (bound %.0f) ms %.2f' % (d['value'], d['e2e']['value'], d['e2e']['h2d_bound_proofs_per_s'], d['ms_per_step']))"
  done
done
