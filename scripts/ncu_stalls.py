import csv, io, subprocess, sys
rep=sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h=rows[1]
ix={n:i for i,n in enumerate(h)}
st=[n for n in h if n.startswith('stall_')]
data=[r for r in rows[2:] if len(r)==len(h) and r[0].startswith('0x')]
def I(r,n):
    try: return int(r[ix[n]] or 0)
    except: return 0
tot=sum(I(r,'# Samples') for r in data); toti=sum(I(r,'Instructions Executed') for r in data)
start=0
segs=[]
for i,r in enumerate(data):
    if 'BAR.SYNC' in r[1] or r[1].strip().startswith('RET') or 'EXIT' in r[1]:
        segs.append((start,i+1)); start=i+1
segs.append((start,len(data)))
for a,b in segs:
    seg=data[a:b]
    ssum=sum(I(r,'# Samples') for r in seg); isum=sum(I(r,'Instructions Executed') for r in seg)
    if ssum/tot<0.005: continue
    agg={n:sum(I(r,n) for r in seg) for n in st}
    tops=sorted(agg.items(), key=lambda kv:-kv[1])[:6]
    print("%6d-%6d samp %5.2f%% inst %5.2f%% | %s"%(a,b,100*ssum/tot,100*isum/toti,' '.join("%s %.0f%%"%(k[6:],100*v/ssum) for k,v in tops)))
if len(sys.argv)>3:
    a,b=int(sys.argv[2]),int(sys.argv[3])
    for r in data[a:b]:
        s=I(r,'# Samples')
        top=sorted(((I(r,n),n[6:]) for n in st),reverse=True)[:2]
        print("%-80s %6d %s"%(r[1][:80], s, ' '.join("%s:%d"%(n,v) for v,n in top if v)))
