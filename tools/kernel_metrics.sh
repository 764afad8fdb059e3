#!/bin/bash
# Per-kernel roofline evidence from ncu for every kernel of one bench step: duration, DRAM bytes read / written, DRAM throughput
# (% of peak), tensor-pipe activity and active warps.  After the same command has run clean without ncu.
# usage (under gpurun): bash tools/kernel_metrics.sh <tag> [workload]   -> gpurun_out/kernel_metrics_<tag>.csv
set -u
TAG=${1:-r2}
WL=${2:-cfg2}
KREGEX='regex:^(umma_gemm|polar_gemm|polar_fused_abm|pooled_eig|jacobi_cluster_global|angles|mix_weights|selector_bwd|selector_corr|importance_rows|split_bf16|pack_bf16|colsum|colsum_reduce|splitk_reduce|importance_mix|mix_teacher|wgrad_dots|wgrad_importance|wgrad_reduce|loss_reduce|polar_prep_student|polar_prep_student_vec|polar_prep_teacher|polar_finish|vt_prep_teacher|vt_augment|vt_theta)_kernel'
CMD="python bench.py --workload $WL --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_km_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread \
    --clock-control none -k "$KREGEX" -c 4000 --csv --log-file gpurun_out/kernel_metrics_$TAG.csv $CMD > gpurun_out/ncu_km_$TAG.log 2>&1
echo "kernel-metrics rc=$?"
