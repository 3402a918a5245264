#!/bin/bash
# bus-kernel A/B: graph-loop time per iteration (no events) and event-bracketed kernel times, several grids
for lib in "$@"; do
  export EXAADMM_B200_LIB=$PWD/$lib
  echo "=== $(basename $lib .so)"
  for w in ACTIVSg70k case13659pegase case2869pegase; do
    EA_KERNEL_TIMING=0 python tools/profile_iter.py $w 100 192 2>&1 | head -1 | cut -c1-75
    python tools/profile_iter.py $w 100 192 2>&1 | head -1 | cut -c1-130
  done
done
