#!/bin/sh
# tuning aid: end-to-end throughput (host buffers) for the chunk-schedule variants P2V_TAPER=0/1/2 and P2V_SPLIT
for t in 0 1 2; do
  P2V_TAPER=$t python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('taper=$t value %.0f e2e %.0f (bound %.0f)' % (d['value'], d['e2e']['value'], d['e2e']['h2d_bound_proofs_per_s']))"
done
