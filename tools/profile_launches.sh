#!/bin/bash
# ncu evidence for one bench step (our kernels only), after the same command has run clean without ncu:
#   1. launch list with device time per launch            -> gpurun_out/launches_<tag>.csv
#   2. DRAM bytes read / written per launch (all kernels) -> gpurun_out/traffic_<tag>.csv
# usage (under gpurun): bash tools/profile_launches.sh <tag>
set -u
TAG=${1:-r1}
KREGEX='regex:^(umma_gemm|polar_gemm|polar_fused_abm|pooled_eig|angles|mix_weights|selector_bwd|importance_rows|split_bf16|pack_bf16|colsum|importance_mix|mix_teacher|wgrad_dots|wgrad_importance|loss_reduce|polar_prep_student|polar_prep_student_vec|polar_prep_teacher|polar_finish)_kernel'
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 4000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$KREGEX" -c 4000 --csv --log-file gpurun_out/traffic_$TAG.csv $CMD > gpurun_out/ncu_traffic_$TAG.log 2>&1
echo "profile rc=$?"
[ -n "${PROFILE_LIGHT:-}" ] && exit 0     # PROFILE_LIGHT=1: launch list + traffic only
#   3. every kernel of the process (torch's included), to account for the step time outside our kernels
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_all_$TAG.csv $CMD > gpurun_out/ncu_all_$TAG.log 2>&1
echo "profile-all rc=$?"
#   4. --set full of the two top kernels (one launch each)
ncu --set full --clock-control none --import-source on -k regex:polar_gemm_kernel -s 68 -c 1 -f -o gpurun_out/polar_gemm_full_$TAG $CMD > gpurun_out/ncu_pg_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pooled_eig_kernel -s 2 -c 1 -f -o gpurun_out/pooled_eig_full_$TAG $CMD > gpurun_out/ncu_pe_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:polar_fused_abm_kernel -s 45 -c 1 -f -o gpurun_out/polar_fused_full_$TAG $CMD > gpurun_out/ncu_pf_$TAG.log 2>&1
echo "profile-full rc=$?"
