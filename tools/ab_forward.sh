#!/bin/bash
# A/B of kernel-selection knobs on the whole step; alternates the configurations twice.  Usage: tools/ab_forward.sh "ENV1" "ENV2" ...
for round in 1 2; do
  for cfg in "$@"; do
    env $cfg python tools/fwd_time.py 30 2>&1 | tail -1
  done
done
