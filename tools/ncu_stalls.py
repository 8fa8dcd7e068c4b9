"""Summarise an `ncu --page source --csv` dump: per kernel, total stall samples by reason and the hottest SASS lines."""
import csv, gzip, sys, collections
path = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 12
which = int(sys.argv[3]) if len(sys.argv) > 3 else None
f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
kern, hdr, rows, out = None, None, [], []
def flush():
    if kern is not None and rows:
        out.append((kern, hdr, list(rows)))
for r in csv.reader(f):
    if len(r) >= 2 and r[0] == "Kernel Name":
        flush(); kern = r[1][:90]; hdr = None; rows = []
    elif len(r) > 5 and r[0] == "Address":
        hdr = r
    elif hdr and len(r) == len(hdr):
        rows.append(r)
flush()
for ki, (k, hdr, rows) in enumerate(out):
    if which is not None and ki != which:
        continue
    si = hdr.index("# Samples")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    for r in rows:
        for i in stall_cols:
            tot[hdr[i]] += int(r[i] or 0)
    n = sum(int(r[si] or 0) for r in rows)
    print(f"== [{ki}] {k}\n   samples {n}: " + ", ".join(f"{a[6:]} {100*b/max(n,1):.0f}%" for a, b in tot.most_common(7)))
    for r in sorted(rows, key=lambda r: -int(r[si] or 0))[:topn]:
        top = max(stall_cols, key=lambda i: int(r[i] or 0))
        print(f"   {int(r[si]):7d} {100*int(r[si])/max(n,1):5.1f}%  {hdr[top][6:]:14s} {r[1].strip()[:90]}")
