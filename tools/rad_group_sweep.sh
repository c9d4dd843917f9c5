#!/bin/bash
# Pair-sweep experiment: time one resident bake of a workload for every column-group size.
W=${1:-config4}
for g in 32 16 8 4; do
  echo "== LTR_RAD_GROUP=$g"
  LTR_RAD_GROUP=$g LTR_TRACE=1 python tools/profile_step.py $W 1 2>&1 | grep -E "pairs-kernel|profiled step" | tail -2
done
