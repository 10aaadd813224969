#!/bin/bash
# ncu --set full captures of round 2 (one GPU). The reports are summarised on the box (tools/ncu_summary.py) and deleted: only the
# text summaries travel back (gpurun merges at most 64 MiB).
NCU="ncu --set full --clock-control none -f"
mkdir -p gpurun_out
cap() { # name, kernel regex, skip, count, target args...
  local name=$1 re=$2 skip=$3 cnt=$4; shift 4
  $NCU -k regex:"$re" --launch-skip $skip -c $cnt -o gpurun_out/$name python tools/profile_target.py "$@" > gpurun_out/ncu_$name.log 2>&1
  { echo "# ncu --set full --clock-control none -k regex:\"$re\" --launch-skip $skip -c $cnt python tools/profile_target.py $*"; python tools/ncu_summary.py gpurun_out/$name.ncu-rep; } > gpurun_out/$name.txt
  rm -f gpurun_out/$name.ncu-rep
}
# C2: the bench kernels (generate+extend, shade, connect, finalize)
cap r2_c2_kernels "wide|shadeKernel|finalizeKernel" 8 4 c2 4
# C3: incoherent bounce rays - extend / shade / connect at path lengths 2 .. 5 of the second frame
cap r2_c3_kernels "wideExtendKernel|shadeKernel|wideConnectKernel" 24 12 c3 3
# C4: two-level traversal of the third frame, then the refit / TLAS kernels
cap r2_c4_traversal "wide" 12 6 c4 3
cap r2_c4_refit "fitKernel|requantKernel|triRewriteKernel|instBoundsKernel|radixTreeKernel|collapseKernel|mortonKernel|triBoundsKernel" 30 14 c4 3
# BLAS build of 1M triangles (PLOC + SAH-optimal collapse) and a refit
cap r2_build_kernels "plocKernel|collapseKernel|fitKernel|mortonKernel|triBoundsKernel|leafBoxKernel|requantKernel|triRewriteKernel" 0 20 build 1
# C5: the filter chain of the third frame
cap r2_c5_filter "prepare|atrous|taaKernel|presentKernel" 16 8 c5 3
ls -la gpurun_out/r2_*.txt
