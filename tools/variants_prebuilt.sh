#!/bin/sh
# run on the GPU box: hash parity (+ full-verifier parity with PARITY=1) + K1 rate + K6a probe for every prebuilt variant (tools/build_variants.py)
for so in plonky2-verifier_b200/variants/libp2v_*.so; do
  echo "== $so"
  P2V_LIB_PATH=$PWD/$so timeout 300 python -m pytest tests/test_gpu_hash.py -m gpu -x -q 2>&1 | tail -1
  if [ -n "$PARITY" ]; then P2V_LIB_PATH=$PWD/$so timeout 600 python -m pytest tests/test_gpu_verify.py -m gpu -x -q 2>&1 | tail -1; fi
  P2V_LIB_PATH=$PWD/$so timeout 300 python tools/perf_poseidon.py 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('K1 perms/s %.4e' % d['perms_per_s'])"
  P2V_LIB_PATH=$PWD/$so timeout 300 python tools/perf_k6a.py ${1:-40000} 2>&1 | tail -1
done
