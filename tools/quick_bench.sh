#!/bin/sh
# tuning helper: one line per variant: value, roofline frac, kernel ms
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --proofs ${1:-50000} 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.0f e2e %.0f frac %.4f fri_merkle %.2f ms challenges %.2f' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['kernel_ms']['fri_merkle'], d['kernel_ms']['challenges']))"
