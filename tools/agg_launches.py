"""Aggregate an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list by
kernel name: total time, launches, DRAM bytes and the DRAM rate they imply.  `--steps N` divides by the number of
training steps the capture holds (tools/train_layers.py runs three)."""
import collections, csv, sys
args = [a for a in sys.argv[1:] if not a.startswith("--")]
steps = 1
for a in sys.argv[1:]:
    if a.startswith("--steps="):
        steps = int(a.split("=")[1])
rows = list(csv.reader(open(args[0])))
hdr = None
per = collections.OrderedDict()      # launch id -> [name, time_ms, bytes]
for r in rows:
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    try:
        v = float(d["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    e = per.setdefault(d["ID"], [d["Kernel Name"].split("(")[0][:70], 0.0, 0.0])
    unit, name = d["Metric Unit"], d["Metric Name"]
    if name.startswith("gpu__time_duration"):
        e[1] += v * {"us": 1e-3, "ns": 1e-6, "s": 1e3, "ms": 1.0, "usecond": 1e-3, "nsecond": 1e-6, "msecond": 1.0, "second": 1e3}.get(unit, 1.0)
    elif name.startswith("dram__bytes"):
        e[2] += v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1.0)
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for name, t, b in per.values():
    a = agg[name]
    a[0] += 1
    a[1] += t
    a[2] += b
tot = sum(v[1] for v in agg.values())
print(f"{'ms/step':>10} {'share':>6} {'launches':>8} {'GB/step':>9} {'TB/s':>6}  kernel")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    rate = v[2] / (v[1] * 1e-3) / 1e12 if v[1] > 0 and v[2] > 0 else 0.0
    print(f"{v[1] / steps:10.3f} {100 * v[1] / tot:5.1f}% {v[0] // steps:8d} {v[2] / steps / 1e9:9.2f} {rate:6.2f}  {k}")
print(f"{tot / steps:10.3f} total")
