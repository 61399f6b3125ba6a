"""Sum warp-instructions and stall samples of an ncu report over named source-line ranges.
usage: python tools/ncu_ranges.py report.ncu-rep units file:lo-hi[=name] ..."""
import csv, subprocess, sys, io
rep = sys.argv[1]; units = float(sys.argv[2])
ranges = []
for a in sys.argv[3:]:
    nm = a
    if "=" in a: a, nm = a.split("=")
    f, r = a.split(":"); lo, hi = r.split("-"); ranges.append((f, int(lo), int(hi), nm))
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur = None; hdr = None; acc = {r[3]: [0, 0] for r in ranges}; tot = [0, 0]
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 6 and r[0] == "Line No": hdr = r; si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed"); continue
    if hdr and len(r) == len(hdr) and r[0] not in ("", "Line No"):
        try: s = int(r[si]); ins = int(r[ii]); ln = int(r[0])
        except ValueError: continue
        tot[0] += ins; tot[1] += s
        for f, lo, hi, nm in ranges:
            if cur == f and lo <= ln <= hi: acc[nm][0] += ins; acc[nm][1] += s
print(f"total {tot[0]/units:.0f} instr/unit, {tot[1]} samples")
for nm, (i, s) in acc.items():
    print(f"{nm:40s} {i/units:9.0f} instr/unit {100*i/tot[0]:5.1f}%   samples {100*s/max(tot[1],1):5.1f}%")
