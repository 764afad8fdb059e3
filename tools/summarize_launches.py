"""Summarise an ncu launch list (gpu__time_duration per launch) by kernel for the LAST bench step."""
import collections, csv, sys
path, steps = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 4
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
names = [(r[ki], float(r[vi].replace(",", "")), r[gi], r[bi]) for r in data if len(r) > vi]
per = len(names) // steps
last = names[-per:]
tot = collections.OrderedDict()
for n, v, g, b in last:
    k = n.split("(")[0]
    k = k.replace("basd::", "")[:90]
    tot.setdefault(k, [0.0, 0]); tot[k][0] += v; tot[k][1] += 1
s = sum(v[0] for v in tot.values())
print(f"# {path}: {len(names)} launches captured, {per} per step; last step, device time by kernel (cold-cache, serialised)")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"{v[0] / 1e6:9.3f} ms {v[1]:4d}x {100 * v[0] / s:6.2f}%  {k}")
print(f"{s / 1e6:9.3f} ms total")
