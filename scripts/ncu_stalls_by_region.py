import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
cur=None; hdr=None
agg=collections.defaultdict(lambda: collections.Counter())
def num(x):
    try: return int(x)
    except: return 0
def region(f,l):
    if f!="conv_halo_tc.cu": return f
    if l<255: return "setup/linear"
    if l<372: return "producer-halo"
    if l<505: return "mma"
    if l<585: return "loader"
    if l<790: return "epilogue"
    return "tail"
for r in rows:
    if len(r)==2 and r[0]=="File Path": cur=r[1].split("/")[-1]; continue
    if r and r[0]=="Line No": hdr=r; idx={h:i for i,h in enumerate(hdr)}; stall=[h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]; continue
    if hdr and r and r[0].isdigit():
        reg=region(cur,int(r[0]))
        for h in stall: agg[reg][h]+=num(r[idx[h]])
        agg[reg]["inst"]+=num(r[idx["Instructions Executed"]])
for reg,c in agg.items():
    tot=sum(v for k,v in c.items() if k!="inst")
    if tot<50: continue
    print(f"{reg:18s} inst {c['inst']:9d} samples {tot:6d} :", ", ".join(f"{k[6:]} {v}" for k,v in c.most_common() if k!='inst' and v>0.03*tot))
