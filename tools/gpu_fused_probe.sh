#!/bin/bash
# Is the fused A/Bm kernel bound per SM or by the chip's HBM while its CTAs load?  bench.py cfg2 with the fused kernel on fewer
# CTAs (BASD_POLAR_FUSED_GRID) and with odd CTAs started late (BASD_POLAR_FUSED_STAGGER, cycles); polar_gemm ms per step each.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 10 --warmup 3 > gpurun_out/fp_$tag.json 2> gpurun_out/fp_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/fp_$tag.json").read().strip().splitlines()[-1])
    print("$tag: step %.3f ms, polar_gemm %.4f ms" % (d["ms_per_step"], d["roofline"]["kernel_ms_per_step"]["polar_gemm"]))
except Exception as e:
    print("$tag unreadable", e)
PY
}
run base BASD_NOOP=1
run grid111 BASD_POLAR_FUSED_GRID=111
run grid74 BASD_POLAR_FUSED_GRID=74
run stag8k BASD_POLAR_FUSED_STAGGER=8000
run stag16k BASD_POLAR_FUSED_STAGGER=16000
run stag24k BASD_POLAR_FUSED_STAGGER=24000
