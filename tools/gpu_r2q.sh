#!/bin/bash
# round-2 session q: parity + A/B of the persistent project / token_gram launches (new default vs BASD_*_TILE=1)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1000 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2q_tests.log
tail -3 gpurun_out/r2q_tests.log
for wl in cfg2 cfg4 cfg5 cfg3; do
  timeout 300 python bench.py --workload $wl --no-cpu-baseline --no-e2e --steps 10 --warmup 3 > gpurun_out/r2q_${wl}_new.json 2> gpurun_out/r2q_${wl}_new.err; echo "$wl new rc=$?"
  BASD_PROJECT_TILE=1 BASD_TOKEN_GRAM_TILE=1 timeout 300 python bench.py --workload $wl --no-cpu-baseline --no-e2e --steps 10 --warmup 3 > gpurun_out/r2q_${wl}_old.json 2> gpurun_out/r2q_${wl}_old.err; echo "$wl old rc=$?"
  python - <<PY
import json
for tag in ("new", "old"):
    try:
        d = json.loads(open("gpurun_out/r2q_${wl}_%s.json" % tag).read().strip().splitlines()[-1])
        k = d["roofline"]["kernel_ms_per_step"]
        print("${wl}", tag, round(d["ms_per_step"], 3), {n: k[n] for n in ("project", "token_gram", "gram", "polar_gemm", "pooled_eig") if n in k})
    except Exception as e:
        print("${wl}", tag, "unreadable", e)
PY
done
