"""Summarise an `ncu --page source --csv` dump: top SASS instructions by stall samples.
usage: ncu -i rep.ncu-rep --page source --csv > src.csv ; python tools/ncu_top.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
tot = sum(int(r[col["# Samples"]]) for r in body)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[col[s]] or 0) for r in body) for s in stalls}
print("total samples", tot, "instructions", len(body))
print("by reason:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
execd = sum(int(r[col["Instructions Executed"]]) for r in body)
print("warp-instructions executed:", execd)
idx = sorted(range(len(body)), key=lambda i: -int(body[i][col["# Samples"]]))[:n]
for i in sorted(idx):
    r = body[i]
    top = sorted(((int(r[col[s]] or 0), s) for s in stalls), reverse=True)[:2]
    print(f"{i:5d} {int(r[col['# Samples']]):6d} {100*int(r[col['# Samples']])/tot:5.1f}%  {r[col['Source']].strip():70s} {top}")
