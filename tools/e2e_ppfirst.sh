#!/bin/sh
# tuning aid: bench.py for the host schedule knobs (transcript-first on/off x launch priorities x tail taper)
for cfg in "0 0 0" "0 2 0" "1 0 0" "1 2 0" "1 2 1056" "1 2 2112" "1 0 1056"; do
  set -- $cfg
  P2V_PPFIRST=$1 P2V_PRIO=$2 P2V_TAIL=$3 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ppfirst=$1 prio=$2 tail=$3 value %.0f e2e %.0f (bound %.0f) ms %.2f' % (d['value'], d['e2e']['value'], d['e2e']['h2d_bound_proofs_per_s'], d['ms_per_step']))"
done
