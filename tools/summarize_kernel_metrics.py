"""Per-kernel summary of tools/kernel_metrics.sh output (last step captured): launches, total time, DRAM GB/s against the
measured copy bandwidth, tensor-pipe activity.  usage: python tools/summarize_kernel_metrics.py <csv> [launches_per_step]"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        rows.append(r)
    # ncu --csv (long format): one row per (launch ID, metric)
    launches = collections.OrderedDict()
    for r in rows:
        lid = int(r["ID"])
        d = launches.setdefault(lid, {"name": r["Kernel Name"]})
        try:
            d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
            d[r["Metric Name"] + ":unit"] = r["Metric Unit"]
        except ValueError:
            pass
    ids = list(launches)
    per_step = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    if not per_step:          # the bench runs 3 warm-up steps + 1 timed + up to 1 breakdown step: take the last full step
        names = [launches[i]["name"] for i in ids]
        first = names[0]
        starts = [k for k, n in enumerate(names) if n == first]
        per_step = starts[1] - starts[0] if len(starts) > 1 else len(names)
    last = ids[-per_step:]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    agg = collections.OrderedDict()
    for i in last:
        d = launches[i]
        name = re.sub(r"\(.*", "", d["name"])
        a = agg.setdefault(name, collections.Counter())
        t_unit = d.get("gpu__time_duration.sum:unit", "ns")
        t = d.get("gpu__time_duration.sum", 0.0) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(t_unit, 1e-9)
        def bytes_of(k):
            u = d.get(k + ":unit", "byte")
            return d.get(k, 0.0) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        a["n"] += 1; a["t"] += t; a["rd"] += bytes_of("dram__bytes_read.sum"); a["wr"] += bytes_of("dram__bytes_write.sum")
        a["tensor_t"] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) * t
        a["dram_t"] += d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0.0) * t
        a["warps_t"] += d.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0.0) * t
        a["regs"] = max(a["regs"], d.get("launch__registers_per_thread", 0.0))
    tot = sum(a["t"] for a in agg.values())
    print(f"# {path}: last step, {per_step} launches, {tot * 1e3:.3f} ms of device time under ncu (cold-cache, serialised: compare shares)")
    print(f"# DRAM GB/s against the measured copy bandwidth {hbm:.0f} GB/s (MEASURED_PEAKS.json); tensor = sm__pipe_tensor_cycles_active % of peak")
    print(f"{'kernel':86s} {'n':>3s} {'ms':>8s} {'share':>6s} {'DRAM MB':>9s} {'GB/s':>7s} {'of copy':>7s} {'dram%':>6s} {'tensor%':>7s} {'warps%':>6s} {'regs':>4s}")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
        gbs = (a["rd"] + a["wr"]) / a["t"] / 1e9 if a["t"] else 0.0
        print(f"{name[:86]:86s} {a['n']:3d} {a['t'] * 1e3:8.3f} {a['t'] / tot:6.1%} {(a['rd'] + a['wr']) / 1e6:9.1f} {gbs:7.0f} {gbs / hbm:7.1%} "
              f"{a['dram_t'] / a['t']:6.1f} {a['tensor_t'] / a['t']:7.1f} {a['warps_t'] / a['t']:6.1f} {int(a['regs']):4d}")


if __name__ == "__main__":
    main()
