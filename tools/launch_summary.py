import csv, collections, sys
rows=[r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h=rows[0]; ik=h.index("Kernel Name"); iv=h.index("Metric Value")
d=collections.defaultdict(list)
for r in rows[1:]:
    n=r[ik].split("(")[0].split("::")[-1][:30]
    d[n].append(float(r[iv].replace(",","")))
for n,v in sorted(d.items(), key=lambda kv:-sum(kv[1]))[:8]:
    v2=sorted(v)
    print(f"{n:32s} n={len(v):5d} total={sum(v)/1e6:9.3f} ms avg={sum(v)/len(v)/1e3:8.2f} us median={v2[len(v2)//2]/1e3:8.2f} max={max(v)/1e3:8.2f}")
