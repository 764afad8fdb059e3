#!/bin/bash
# compute-sanitizer evidence for the shared-memory / cluster / mbarrier kernels (pooled_eig, angles, polar_fused_abm, polar_gemm):
# memcheck and racecheck over one small forward + backward (smoke(): cfg1 shapes at B = 4 -> D_s = 192, N = 196: the cluster
# Jacobi with st.async mailboxes, the fused polar kernel and every tcgen05 GEMM variant of the feature form run).
# usage (under gpurun): bash tools/sanitize.sh <tag>   -> gpurun_out/sanitizer_<tool>_<tag>.log
set -u
TAG=${1:-r2}
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_${tool}_$TAG.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/sanitizer_${tool}_$TAG.log
  tail -4 gpurun_out/sanitizer_${tool}_$TAG.log
done
