import csv, re, sys
lines=[l for l in open(sys.argv[1]) if not l.startswith("==")]
rows=list(csv.DictReader(lines))
idx=[i for i,r in enumerate(rows) if "compact_mask" in r["Kernel Name"]]
k=int(sys.argv[2]) if len(sys.argv)>2 else 2
a,b=idx[k],idx[k+1]
tot=0
for r in rows[a:b]:
    v=float(r["Metric Value"].replace(",",""))/(1000 if r["Metric Unit"]=="ns" else 1)
    tot+=v
    name=re.sub(r"\(.*","",r["Kernel Name"]).replace("void ","").replace("<unnamed>::","")[:64]
    print(f"{v:8.2f} {r['Grid Size']:>16s} {r['Block Size']:>13s} {name}")
print("frame total us", round(tot,1), "kernels", b-a)
