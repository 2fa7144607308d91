import csv, sys, re, collections
f=sys.argv[1]
rows=list(csv.reader(open(f)))
hdr=rows[1]; data=rows[2:]
ia=hdr.index("Source"); ie=hdr.index("Instructions Executed"); isamp=hdr.index("# Samples")
tot_e=sum(int(r[ie]) for r in data if r[ie].isdigit()); tot_s=sum(int(r[isamp]) for r in data if r[isamp].isdigit())
ex=collections.Counter(); sm=collections.Counter()
for r in data:
    s=r[ia].strip()
    m=re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_]+)',s)
    op=m.group(2) if m else s[:10]
    ex[op]+=int(r[ie]); sm[op]+=int(r[isamp])
print("total warp-inst",tot_e,"samples",tot_s)
for op,c in ex.most_common(int(sys.argv[2]) if len(sys.argv)>2 else 25):
    print("%-12s exec %6.2f%%  samples %6.2f%%"%(op,100*c/tot_e,100*sm[op]/tot_s))
# stall reason columns by top instructions
if len(sys.argv)>3:
    cols=[i for i,h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    top=sorted(data,key=lambda r:-int(r[isamp]))[:int(sys.argv[3])]
    for r in top:
        st={hdr[i][6:]:int(r[i]) for i in cols if r[i].isdigit() and int(r[i])>0}
        print(r[isamp],r[ie],r[ia].strip()[:60],st)
