#!/bin/bash
# A/B timing of decode variants (CNL_LIB selects another build of the same library).  Usage: tools/ab_decode.sh out_prefix
out=${1:-gpurun_out/ab_decode}
for lib in "" $(ls centernet-lightning_b200/ab/*.so 2>/dev/null); do
  name=$(basename "${lib:-default}" .so)
  if [ -n "$lib" ]; then export CNL_LIB=$PWD/$lib; else unset CNL_LIB; fi
  timeout 120 python tools/bench_decode.py --net > ${out}_${name}.log 2>&1
  echo "== $name"; head -6 ${out}_${name}.log | cut -c1-200
done
