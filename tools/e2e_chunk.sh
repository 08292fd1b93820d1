#!/bin/sh
# tuning aid: end-to-end throughput of bench.py against the host chunk size, the schedule (P2V_PPFIRST) and launch priorities (P2V_PRIO)
for c in ${CHUNKS:-0 1056 2112 3200 4256 6400}; do
  set -- $c
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-chunk $1 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('chunk=$1 value %.0f e2e %.0f (bound %.0f) ms %.2f frac %.4f' % (d['value'], d['e2e']['value'], d['e2e']['h2d_bound_proofs_per_s'], d['ms_per_step'], d['roofline']['frac']))"
done
