import csv, sys
rows=[r for r in csv.DictReader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
names=[(r["Kernel Name"], float(r["Metric Value"])/1e3) for r in rows]
idx=[i for i,(n,_) in enumerate(names) if "sample_latent" in n]
i0,i1=idx[20],idx[21]
tot=0
for n,v in names[i0+1:i1+1]:
    print(f"{n[:62]:62s} {v:8.1f}")
    tot+=v
print("step total", tot)
