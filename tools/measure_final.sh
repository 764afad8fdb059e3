#!/bin/bash
# The measurement set behind DESIGN.md section 6 / profiles/<tag>_*: every BASELINE configuration, the views variant, the reference
# arm, then the ncu launch list, DRAM traffic and per-kernel metrics of one cfg2 step (each after the same command ran clean).
# usage (under gpurun): bash tools/measure_final.sh <tag>
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
TAG=${1:-r2x}
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/${TAG}_bench_cfg2.json 2> gpurun_out/${TAG}_bench_cfg2.err; echo "cfg2 rc=$?"
timeout 300 python bench.py --views --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg2_views.json 2> gpurun_out/${TAG}_bench_cfg2_views.err; echo "views rc=$?"
for wl in cfg1 cfg3 cfg4 cfg5; do
  timeout 400 python bench.py --workload $wl --no-cpu-baseline > gpurun_out/${TAG}_bench_${wl}.json 2> gpurun_out/${TAG}_bench_${wl}.err; echo "$wl rc=$?"
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "reference rc=$?"
PROFILE_LIGHT=1 bash tools/profile_launches.sh $TAG
bash tools/kernel_metrics.sh $TAG
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d.get("ms_per_step", 0), 3), round(d.get("value", 0), 1), (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
