import csv, io, subprocess, sys, collections
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur=None; hdr=None
agg=collections.defaultdict(lambda: [0,0,0,""])
for r in rows:
    if len(r)==2 and r[0]=="File Path": cur=r[1].split("/")[-1]; continue
    if len(r)>10 and r[1]=="Source" and r[0] in ("Line No","#"): hdr=r; continue
    if hdr and cur and len(r)==len(hdr) and r[0].isdigit():
        ix={h:i for i,h in enumerate(hdr)}
        def g(k):
            try: return int(r[ix[k]] or 0)
            except: return 0
        a=agg[(cur,int(r[0]))]
        a[0]+=g("L2 Theoretical Sectors Global"); a[1]+=g("L2 Theoretical Sectors Local"); a[2]+=g("L1 Tag Requests Global"); a[3]=r[1].strip()[:100]
tot=sum(a[0] for a in agg.values()); totl=sum(a[1] for a in agg.values())
print("total L2 theoretical sectors global: %d (%.1f GB), local: %d (%.1f GB)" % (tot, tot*32/1e9, totl, totl*32/1e9))
for (f,l),a in sorted(agg.items(), key=lambda kv:-kv[1][int(sys.argv[2]) if len(sys.argv)>2 else 0])[:28]:
    print("%9.2f GB global %9.2f GB local  %s:%d  %s" % (a[0]*32/1e9, a[1]*32/1e9, f, l, a[3]))
