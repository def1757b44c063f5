import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5 and r[0].isdigit()]
idx=[i for i,r in enumerate(rows) if 'conv_in' in r[4]]
start=idx[0]
seq=rows[start:start+26]
gf=[0.074,4.756,None,2.359,4.719,None,2.340,4.681,None,2.340,4.681,None,2.265,4.530,1.007,9.362,4.681,1.040,9.362,4.681,1.040,9.437,4.719,1.049,9.512,4.756]
names=['inc.0','inc.3','pool','d1.0','d1.3','pool','d2.0','d2.3','pool','d3.0','d3.3','pool','d4.0','d4.3','up1','u1.0','u1.3','up2','u2.0','u2.3','up3','u3.0','u3.3','up4','u4.0','u4.3+outc']
B=int(sys.argv[2]) if len(sys.argv)>2 else 32
tot=0
for n,g,r in zip(names,gf,seq):
    t=float(r[-1])/1e6; tot+=t
    print(f"{n:10s} grid {r[8]:16s} {t:7.3f} ms", f"{g*B/t:7.0f} TF/s" if g else "")
print("total ms", tot, "TF/s", 93.4*B/tot)
