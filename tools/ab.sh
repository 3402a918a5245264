#!/bin/bash
# A/B of library builds on the same box: tools/ab.sh out_prefix lib1 lib2 ...  (paths relative to the repo root)
# per build: lone-lane probe on the hard-branch fixture, three iteration windows of the 70k solve, whole-solve trace.
out=$1; shift
mkdir -p gpurun_out
for lib in "$@"; do
  tag=$(basename $lib .so)
  export EXAADMM_B200_LIB=$PWD/$lib
  {
    echo "=== $tag"
    python tools/probe_branches.py
    python tools/profile_iter.py ACTIVSg70k 5 20 | grep -v "^{"
    python tools/profile_iter.py ACTIVSg70k 40 40 | grep -v "^{"
    python tools/profile_iter.py ACTIVSg70k 100 200 | grep -v "^{"
    python tools/profile_iter.py case13659pegase 40 100 | grep -v "^{"
    python tools/iter_trace.py ACTIVSg70k 2>&1 >/dev/null | grep "^#"
  } > gpurun_out/${out}_${tag}.txt 2>&1
done
cat gpurun_out/${out}_*.txt
