import csv, subprocess, sys, io, collections
rep=sys.argv[1]; launch=sys.argv[2]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--launch-skip", launch, "--launch-count", "1"], capture_output=True, text=True).stdout
cur=None; out=[]
for r in csv.reader(io.StringIO(src)):
    if len(r)==2 and r[0]=="File Path": cur=r[1].split("/")[-1]; continue
    if len(r)>7 and r[0].isdigit():
        try: out.append((int(r[7]), int(r[6]), cur, int(r[0]), r[1].strip()[:90]))
        except ValueError: pass
tot=sum(o[0] for o in out); ts=sum(o[1] for o in out)
# group by region of conv_halo_tc.cu
reg=collections.Counter(); regs=collections.Counter()
def region(f,l):
    if f!="conv_halo_tc.cu": return f
    if l<140: return "setup"
    if l<230: return "producer-linear"
    if l<300: return "producer-halo"
    if l<430: return "mma"
    if l<500: return "loader"
    if l<680: return "epilogue"
    return "tail"
for n,s,f,l,t in out:
    reg[region(f,l)]+=n; regs[region(f,l)]+=s
print("total inst",tot,"samples",ts)
for k,v in reg.most_common(): print(f"{k:28s} inst {v:10d} {100*v/tot:5.1f}%  samples {regs[k]:6d} {100*regs[k]/ts:5.1f}%")
print("top by inst")
for n,s,f,l,t in sorted(out,key=lambda o:-o[0])[:int(sys.argv[3]) if len(sys.argv)>3 else 25]:
    print(f"  inst {n:9d} {100*n/tot:5.1f}%  smp {s:5d}  {f}:{l}  {t}")
