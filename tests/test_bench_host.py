"""Host-side pieces of bench.py that do not need a GPU: the workload table against BASELINE.json's configurations (SURVEY.md
section 8a) and the NUMA binding helper's behaviour on a box without NVML."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def test_workloads_are_the_baseline_configurations():
    # B, N_s / N_t, D_s / D_t, L_t / H per configuration (cfg5: the per-GPU shard of the global batch 1024)
    want = {"cfg1": (32, 196, 196, 192, 384, 12, 6), "cfg2": (256, 196, 196, 192, 768, 12, 12), "cfg3": (256, 196, 49, 384, 2048, 1, 1),
            "cfg4": (256, 196, 196, 384, 1024, 24, 16), "cfg5": (128, 576, 576, 384, 768, 12, 12)}
    for name, (B, Ns, Nt, Ds, Dt, Lt, H) in want.items():
        w = bench.workload(0, name)
        assert (w.B, w.Ns, w.Nt, w.Ds, w.Dt, w.Lt, w.H) == (B, Ns, Nt, Ds, Dt, Lt, H), name
    assert bench.workload(64, "cfg2").B == 64                      # --batch overrides the per-GPU batch (weak scaling)
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert len(base["configs"]) == len(want)


def test_numa_binding_is_optional():
    """Without a driver / NVML the helper returns None and leaves the affinity alone (the bench then runs unbound)."""
    before = os.sched_getaffinity(0)
    n = bench.bind_to_gpu_numa_node(0)
    assert n is None or (isinstance(n, int) and n >= 1)
    if n is None:
        assert os.sched_getaffinity(0) == before
    else:
        os.sched_setaffinity(0, before)
