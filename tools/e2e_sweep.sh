#!/bin/sh
for c in 4096 8192 16384 32768; do
  printf "chunk %s : " $c
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-chunk $c 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.0f e2e %.0f' % (d['value'], d['e2e']['value']))"
done
