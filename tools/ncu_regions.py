"""Summarise an ncu report by source region (20-line buckets): instructions and stall samples.
usage: python tools/ncu_regions.py report.ncu-rep units_per_launch [top]"""
import csv, subprocess, io, collections, sys
rep = sys.argv[1]; units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
txt = subprocess.run(["ncu","-i",rep,"--page","source","--print-source","cuda,sass","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(txt)))
cur=None;hdr=None;agg=collections.Counter();samp=collections.Counter()
for r in rows:
    if len(r)>=2 and r[0]=="File Path": cur=r[1].split("/")[-1];continue
    if len(r)>6 and r[0]=="Line No": hdr=r;si=hdr.index("# Samples");ii=hdr.index("Instructions Executed");continue
    if hdr and len(r)==len(hdr) and r[0] not in("","Line No"):
        try: s=int(r[si]);ins=int(r[ii])
        except ValueError: continue
        k=(cur,int(r[0])//20*20); agg[k]+=ins; samp[k]+=s
tot=sum(agg.values()); ts=sum(samp.values())
print(f"total warp-instructions {tot} = {tot/units:.0f} per unit; samples {ts}")
for k,v in sorted(agg.items(), key=lambda kv:-kv[1])[:top]:
    print(f"{k[0]}:{k[1]:<5} instr {100*v/tot:5.1f}%  samples {100*samp[k]/max(ts,1):5.1f}%   {v/units:.0f}/unit")
