#!/bin/sh
# tuning aid: end-to-end throughput of bench.py against the host chunk size and the tail taper (transcript-first schedule)
for cfg in ${CFGS:-"0 0" "1056 0" "2112 0" "2112 1056" "3200 0" "3200 1056" "4256 2112" "6400 0" "6400 1056"}; do
  set -- $cfg
  P2V_TAIL=$2 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-chunk $1 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('chunk=$1 tail=$2 value %.0f e2e %.0f (bound %.0f) ms %.2f frac %.4f' % (d['value'], d['e2e']['value'], d['e2e']['h2d_bound_proofs_per_s'], d['ms_per_step'], d['roofline']['frac']))"
done
