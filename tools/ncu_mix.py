"""Opcode mix (executed warp-instructions per unit) and per-region profile from an .ncu-rep source page.
usage: python tools/ncu_mix.py rep.ncu-rep units [window]"""
import csv, collections, subprocess, sys, io
rep, units = sys.argv[1], float(sys.argv[2])
win = int(sys.argv[3]) if len(sys.argv) > 3 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; col = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
def opc(r):
    toks = r[col["Source"]].strip().split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    parts = op.split(".")
    return parts[0] + ("." + parts[1] if parts[0] in ("LDS", "STS", "LDG", "STG", "LDL", "STL") and len(parts) > 1 else "")
cnt = collections.Counter(); smp = collections.Counter()
for r in body:
    cnt[opc(r)] += int(r[col["Instructions Executed"]]); smp[opc(r)] += int(r[col["# Samples"]])
tot = sum(cnt.values())
print(f"executed per unit {tot / units:.1f}; static instructions {len(body)}; samples {sum(smp.values())}")
print("  ".join(f"{op} {c / units:.1f}" for op, c in cnt.most_common(28)))
if win:
    for s in range(0, len(body), win):
        w = body[s:s + win]
        ex = sum(int(r[col["Instructions Executed"]]) for r in w) / units
        sm = sum(int(r[col["# Samples"]]) for r in w)
        ops = collections.Counter(opc(r) for r in w)
        print(f"{s:5d} exec/unit {ex:7.1f} samples {sm:5d}  " + " ".join(f"{o}:{c}" for o, c in ops.most_common(7)))
