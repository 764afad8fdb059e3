"""Summarise ncu CSV launch lists by kernel for the LAST bench step.
    python tools/summarize_launches.py launches.csv [steps]                 device time per kernel
    python tools/summarize_launches.py --traffic traffic.csv [steps] out.json   DRAM bytes per launch per kernel -> json"""
import collections, csv, json, sys

SLOT = {"polar_gemm_kernel": "polar_gemm", "polar_fused_abm_kernel": "polar_gemm", "pooled_eig_kernel": "pooled_eig", "angles_kernel": "angles", "wgrad_dots_kernel": "wgrad_dots",
        "mix_teacher_kernel": "mix_teacher", "colsum_kernel": "colsum", "polar_prep_student_kernel": "polar_prep", "polar_prep_student_vec_kernel": "polar_prep",
        "polar_prep_teacher_kernel": "polar_prep", "polar_finish_kernel": "polar_finish", "importance_rows_kernel": "importance_rows"}


def load(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    return hdr, [r for r in data if len(r) == len(hdr)]


def short(name):
    k = name.split("(")[0].replace("basd::", "").replace("void ", "")
    if "unnamed>::" in k:
        k = k.split("unnamed>::")[1]
    return k[:100]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_ns(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(unit, 1)


if sys.argv[1] == "--traffic":
    path, steps, out = sys.argv[2], int(sys.argv[3]), sys.argv[4]
    hdr, data = load(path)
    idi, ki, mi, ui, vi = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    launches = collections.OrderedDict()
    for r in data:
        d = launches.setdefault(r[idi], {"name": short(r[ki])})
        if r[mi].startswith("dram__bytes"):
            d[r[mi]] = to_bytes(r[vi], r[ui])
        else:
            d["ns"] = to_ns(r[vi], r[ui])
    ls = list(launches.values())
    last = ls[-(len(ls) // steps):]
    agg = collections.OrderedDict()
    for d in last:
        base = d["name"].split("<")[0]
        a = agg.setdefault(SLOT.get(base, d["name"]), {"launches": 0, "dram_read": 0.0, "dram_write": 0.0, "ns": 0.0})
        a["launches"] += 1; a["dram_read"] += d.get("dram__bytes_read.sum", 0); a["dram_write"] += d.get("dram__bytes_write.sum", 0); a["ns"] += d.get("ns", 0)
    res = {k: {"launches_per_step": a["launches"], "dram_bytes_per_launch": (a["dram_read"] + a["dram_write"]) / a["launches"],
               "dram_read_bytes_per_launch": a["dram_read"] / a["launches"], "dram_write_bytes_per_launch": a["dram_write"] / a["launches"],
               "ncu_time_us_per_launch": a["ns"] / a["launches"] / 1e3} for k, a in agg.items()}
    res["_source"] = f"ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, last of {steps} bench steps ({path})"
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))
    sys.exit(0)

path, steps = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 4
hdr, data = load(path)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
names = [(short(r[ki]), to_ns(r[vi], r[ui])) for r in data]
per = len(names) // steps
last = names[-per:]
tot = collections.OrderedDict()
for n, v in last:
    tot.setdefault(n, [0.0, 0]); tot[n][0] += v; tot[n][1] += 1
s = sum(v[0] for v in tot.values())
print(f"# {path}: {len(names)} launches captured, {per} per step; last step, device time by kernel (cold-cache, serialised)")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"{v[0] / 1e6:9.3f} ms {v[1]:4d}x {100 * v[0] / s:6.2f}%  {k}")
print(f"{s / 1e6:9.3f} ms total")
