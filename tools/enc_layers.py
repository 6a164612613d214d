import csv, sys
sys.path.insert(0, '.')
from oracle.aa_oracle import encoder_layer_table
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
T = encoder_layer_table(); tot = 0; L = 131072
agg = {}
for i, (r, ly) in enumerate(zip(rows[1:], T)):
    us = float(r[vi].replace(',', '')) / 1e3
    lout = (L + 2 * ly['pad'] - ly['dil'] * (ly['k'] - 1) - 1) // ly['stride'] + 1
    fl = 2 * B * ly['cin'] * ly['cout'] * ly['k'] * lout
    byt = B * (L * ly['cin'] + lout * ly['cout'] * (2 if ly['res'] == 'end' else 1)) * 2
    print(f"{i:2d} cin {ly['cin']:4d} cout {ly['cout']:4d} k{ly['k']} s{ly['stride']} L{L:6d} {us:8.1f} us {fl/us/1e6:7.1f} TF/s {byt/us/1e3:7.1f} GB/s")
    tot += us; L = lout
print("total us", tot, " -> TF/s", 68.17e9 * B / tot / 1e6)
