#!/bin/sh
# ncu --set full of the Merkle kernel for every prebuilt variant (one launch each); reports land in gpurun_out/
for so in plonky2-verifier_b200/variants/libp2v_*.so; do
  tag=$(basename $so .so | sed s/libp2v_//)
  P2V_LIB_PATH=$PWD/$so timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_fri_merkle -s 2 -c 1 -f -o gpurun_out/r2_$tag python tools/perf_k6a.py ${1:-8192} > gpurun_out/ncu_$tag.log 2>&1
  tail -2 gpurun_out/ncu_$tag.log
done
