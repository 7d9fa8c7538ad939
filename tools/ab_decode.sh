#!/bin/bash
# A/B timing of decode variants (CNL_LIB selects another build of the same library).  Usage: tools/ab_decode.sh out_prefix
out=${1:-gpurun_out/ab_decode}
for lib in "" centernet-lightning_b200/ab/libcnl_nt0.so centernet-lightning_b200/ab/libcnl_nt1.so; do
  name=$(basename "${lib:-default}" .so)
  if [ -n "$lib" ]; then export CNL_LIB=$PWD/$lib; else unset CNL_LIB; fi
  timeout 120 python tools/bench_decode.py --net > ${out}_${name}.log 2>&1
  echo "== $name"; head -4 ${out}_${name}.log
done
