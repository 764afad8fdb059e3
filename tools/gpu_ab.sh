#!/bin/bash
# A/B of one build under an environment switch:  tools/gpu_ab.sh <tag> "<ENV=1 ...>" <workloads...>
# runs the GPU parity tests, then bench.py per workload with and without the switch; per-kernel ms side by side
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
tag=$1; envs=$2; shift 2
mkdir -p gpurun_out
timeout 1000 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${tag}_tests.log
tail -3 gpurun_out/${tag}_tests.log
for wl in "$@"; do
  timeout 300 python bench.py --workload $wl --no-cpu-baseline --no-e2e --steps 10 --warmup 3 > gpurun_out/${tag}_${wl}_new.json 2> gpurun_out/${tag}_${wl}_new.err; echo "$wl new rc=$?"
  env $envs timeout 300 python bench.py --workload $wl --no-cpu-baseline --no-e2e --steps 10 --warmup 3 > gpurun_out/${tag}_${wl}_old.json 2> gpurun_out/${tag}_${wl}_old.err; echo "$wl old rc=$?"
  python - <<PY
import json
r = {}
for t in ("new", "old"):
    try:
        r[t] = json.loads(open("gpurun_out/${tag}_${wl}_%s.json" % t).read().strip().splitlines()[-1])
    except Exception as e:
        print("${wl}", t, "unreadable", e)
if len(r) == 2:
    kn, ko = r["new"]["roofline"]["kernel_ms_per_step"], r["old"]["roofline"]["kernel_ms_per_step"]
    print("${wl}: step", round(r["old"]["ms_per_step"], 3), "->", round(r["new"]["ms_per_step"], 3), "ms;",
          "; ".join(f"{k} {ko.get(k)} -> {kn[k]}" for k in kn if abs(kn[k] - ko.get(k, 0)) > 0.01))
PY
done
