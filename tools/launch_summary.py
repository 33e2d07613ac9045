"""summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        h, start = r, i + 1
        break
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
ui = h.index("Metric Unit")
d = collections.defaultdict(lambda: [0, 0.0])
for r in rows[start:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
    k = r[ki].replace("<unnamed>::", "").replace("void ", "")
    d[k[:80]][0] += 1
    d[k[:80]][1] += v * scale
tot = sum(v[1] for v in d.values())
print(f"{'ms':>10s} {'launches':>8s} {'share':>6s}  kernel   (total {tot:.3f} ms)")
for k, v in sorted(d.items(), key=lambda x: -x[1][1]):
    print(f"{v[1]:10.3f} {v[0]:8d} {100 * v[1] / tot:5.1f}%  {k}")
