"""Executed warp instructions of one kernel aggregated over SASS regions.  usage: ncu_regions.py report kernel_regex [bucket]"""
import csv, subprocess, io, sys
rep, rx = sys.argv[1], sys.argv[2]
B = int(sys.argv[3]) if len(sys.argv) > 3 else 100
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--kernel-name","regex:"+rx],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
start=next(i for i,r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr=rows[start]
ei,src,si=hdr.index("Instructions Executed"),hdr.index("Source"),hdr.index("Warp Stall Sampling (All Samples)")
data=[]
for r in rows[start+1:]:
    if r==hdr: break
    try: data.append((int(r[ei]),int(r[si]),r[src]))
    except: pass
tot=sum(d[0] for d in data); st=sum(d[1] for d in data)
print("total executed", tot, "stall samples", st, "SASS instructions", len(data))
for b in range(0,len(data),B):
    seg=data[b:b+B]
    e=sum(d[0] for d in seg); s=sum(d[1] for d in seg)
    ops={}
    for d in seg:
        t=d[2].split()
        op=t[1] if t[0].startswith('@') else t[0]
        ops[op]=ops.get(op,0)+d[0]
    top=sorted(ops.items(),key=lambda x:-x[1])[:5]
    print("%5d-%5d exec %8d %5.1f%% stalls %5.1f%%  %s"%(b,b+B,e,100*e/tot,100*s/max(st,1),top))
