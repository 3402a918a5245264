#!/bin/bash
# quick A/B: tools/ab_quick.sh out_prefix lib1 lib2 ...  - iteration windows of the 70k solve + smaller grids
out=$1; shift
for lib in "$@"; do
  tag=$(basename $lib .so)
  export EXAADMM_B200_LIB=$PWD/$lib
  {
    echo "=== $tag"
    python tools/profile_iter.py ACTIVSg70k 5 20 | grep -v "^{"
    python tools/profile_iter.py ACTIVSg70k 100 200 | grep -v "^{"
    python tools/profile_iter.py case13659pegase 40 100 | grep -v "^{"
    python tools/profile_iter.py case2869pegase 40 100 | grep -v "^{"
  } > gpurun_out/${out}_${tag}.txt 2>&1
done
cat gpurun_out/${out}_*.txt
