import csv,re,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
tot=0; out=[]
for r in rows[1:]:
    n=re.sub(r'\(.*','',r[ki]); us=float(r[vi].replace(',',''))/1e3
    if us>20 and 'uniform' not in n: out.append(f"{us:.0f}{'*' if '64, 1' in n else ''}"); tot+=us
print(' '.join(out)); print(tot)
