"""SASS opcode census of libaa_b200.so per kernel: the Blackwell-native mnemonics (UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,
UTMALDG/UBLKCP = TMA, SYNCS = mbarrier) plus the packed fp32x2 ops (FFMA2/FADD2/FMUL2).  usage: python tools/sass_census.py > profiles/sass_census_rNN.txt"""
import collections
import os
import re
import subprocess
import sys

so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "audio-algebra_b200", "libaa_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = re.compile(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)")
want = re.compile(r"^(UTC[A-Z]*MMA|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|SYNCS|FFMA2|FADD2|FMUL2|HMMA|UTCBAR|ELECT)")
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in sass.splitlines():
    if line.lstrip().startswith("Function :"):
        cur = subprocess.run(["cu++filt", line.split(":", 1)[1].strip()], capture_output=True, text=True).stdout.strip()
        cur = cur.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        depth, cut = 0, len(cur)
        for i, ch in enumerate(cur):          # drop the parameter list: first '(' outside the template brackets
            if ch == "<":
                depth += 1
            elif ch == ">":
                depth -= 1
            elif ch == "(" and depth == 0:
                cut = i
                break
        cur = cur[:cut].replace("void ", "")
        counts[cur] = collections.Counter()
        continue
    m = pat.match(line)
    if m and cur:
        counts[cur]["_all"] += 1
        w = want.match(m.group(1))
        if w:
            counts[cur][w.group(1)] += 1
            total[w.group(1)] += 1
print(f"# {os.path.basename(so)}: {len(counts)} kernels, sm_100a SASS (cuobjdump -sass); columns = instruction counts in the kernel's SASS")
print("# totals:", ", ".join(f"{k} {v}" for k, v in sorted(total.items())))
for k, c in counts.items():
    feats = ", ".join(f"{n} {v}" for n, v in sorted(c.items()) if n != "_all")
    print(f"{k:70s} {c['_all']:6d} instr  {feats}")
