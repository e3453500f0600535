"""Summarise an ncu --csv launch list (gpu__time_duration.sum) by kernel name (+ grid): count, total, mean."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    name = re.sub(r"^void ", "", name)
    key = (name[:70], row.get("Grid Size", ""))
    v = float(row["Metric Value"].replace(",", ""))
    unit = row.get("Metric Unit", "ns")
    v = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    agg[key][0] += 1; agg[key][1] += v; tot += v
print(f"total {tot/1e3:.2f} ms over {sum(a[0] for a in agg.values())} launches")
byname = collections.defaultdict(float)
for (n, g), (c, t) in agg.items():
    byname[n] += t
for n, t in sorted(byname.items(), key=lambda x: -x[1]):
    print(f"  {t/1e3:8.2f} ms {100*t/tot:5.1f}%  {n}")
print("-- by (kernel, grid)")
for (n, g), (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"  {t/1e3:8.2f} ms  n={c:4d}  mean {t/c:8.1f} us  grid {g:>18s}  {n}")
