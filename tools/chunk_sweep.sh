#!/bin/sh
# host-buffer (e2e) throughput for constant chunk sizes (no ramp below 8192) x pipeline depth
for c in 2048 3072 4096 6144; do for d in 3 4; do
  printf "chunk %s depth %s: " $c $d; python tools/e2e_probe.py 100000 $c $d 2>&1 | grep "^step" | tail -3 | awk '{s+=$5} END {print s/3, "proofs/s"}'
done; done
