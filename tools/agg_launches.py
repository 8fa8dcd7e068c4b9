"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    name = d["Kernel Name"].split("(")[0][:70]
    try:
        v = float(d["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    v *= {"us": 1e-3, "ns": 1e-6, "s": 1e3, "ms": 1.0}.get(d["Metric Unit"], 1.0)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{'ms':>10} {'share':>6} {'launches':>8}  kernel")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:10.3f} {100 * v[1] / tot:5.1f}% {v[0]:8d}  {k}")
print(f"{tot:10.3f} total")
