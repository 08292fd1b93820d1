#!/bin/sh
# correctness + K1 throughput + full bench for compile-time variants
for v in "$@"; do
  P2V_EXTRA_NVCC="$v" python plonky2-verifier_b200/build.py > /dev/null 2>&1 || { echo "build failed: $v"; continue; }
  echo "== $v"
  python -m pytest tests/test_gpu_hash.py -m gpu -x -q 2>&1 | tail -1
  python tools/perf_poseidon.py 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('K1 perms/s %.4e' % d['perms_per_s'])"
  tools/quick_bench.sh 50000
done
