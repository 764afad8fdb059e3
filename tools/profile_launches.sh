#!/bin/bash
# ncu evidence for one bench step (our kernels only), after the same command has run clean without ncu:
#   1. launch list with device time per launch            -> gpurun_out/launches_<tag>.csv
#   2. DRAM bytes read / written per launch (all kernels) -> gpurun_out/traffic_<tag>.csv
# usage (under gpurun): bash tools/profile_launches.sh <tag>
set -u
TAG=${1:-r1}
KREGEX='regex:^(umma_gemm|polar_gemm|pooled_eig|angles|mix_weights|selector_bwd|importance_rows|split_bf16|pack_bf16|colsum|importance_mix|mix_teacher|wgrad_dots|wgrad_importance|loss_reduce|polar_prep_student|polar_prep_teacher|polar_finish)_kernel'
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 4000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$KREGEX" -c 4000 --csv --log-file gpurun_out/traffic_$TAG.csv $CMD > gpurun_out/ncu_traffic_$TAG.log 2>&1
echo "profile rc=$?"
