"""DRAM traffic of the conv_fwd_kernel launches of ONE training step, from an
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:conv_fwd_kernel --csv` log.
Writes profiles/<round>_conv_traffic.json (bench.py reports its per-launch mean as roofline.traffic):
    python tools/conv_traffic.py <ncu csv> [r02]"""
import collections, csv, json, os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if len(r) > 5 and r[0] == "ID")
ki, mi, ui, vi, ii = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "ID"))
per = collections.defaultdict(dict)
for r in rows:
    if len(r) != len(hdr) or r[0] == "ID":
        continue
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1)
    per[r[ii]][r[mi]] = v
n = len(per)
rd = sum(d.get("dram__bytes_read.sum", 0) for d in per.values())
wr = sum(d.get("dram__bytes_write.sum", 0) for d in per.values())
ms = sum(d.get("gpu__time_duration.sum", 0) for d in per.values())
# algorithmic bytes of the same launches (conv inputs 205 M + outputs 182 M elements / image, bf16,
# forward; dgrad reads dY and writes dX = the same two tensors), bs = 32
alg = 2 * (205e6 + 182e6) * 2 * 32
out = dict(launches=n, dram_read_bytes=rd, dram_write_bytes=wr, dram_bytes_per_launch=(rd + wr) / max(n, 1),
           serialized_ms=ms, algorithmic_bytes=alg, algorithmic_bytes_per_launch=alg / max(n, 1),
           note="cold-cache ncu pass over the conv_fwd_kernel launches (forward + dgrad) of one yolov4 bs=32 800x800 step")
print(json.dumps(out, indent=1))
json.dump(out, open(os.path.join(ROOT, "profiles", (sys.argv[2] if len(sys.argv) > 2 else "r02") + "_conv_traffic.json"), "w"),
          indent=1)
