"""Summarise an ncu report per CUDA source line: stall samples and instructions.
usage: python tools/ncu_lines.py report.ncu-rep [top]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file = None; hdr = None; out = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if len(r) > 6 and r[0] == "Line No":
        hdr = r; si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed")
        names = hdr; continue
    if hdr and len(r) == len(hdr) and r[0] not in ("", "Line No"):
        try:
            s = int(r[si]); ins = int(r[ii])
        except ValueError:
            continue
        stalls = {}
        for k, name in enumerate(names):
            if name.startswith("stall_") and "Not Issued" not in name:
                try:
                    v = int(r[k])
                except ValueError:
                    v = 0
                if v:
                    stalls[name[6:]] = v
        out.append((s, ins, cur_file, r[0], r[1].strip()[:90], stalls))
tot = sum(o[0] for o in out)
print("total samples", tot, "total warp-instructions", sum(o[1] for o in out))
for s, ins, f, ln, src, st in sorted(out, reverse=True)[:top]:
    tops = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"{100*s/max(tot,1):5.1f}% {ins:>10} {f}:{ln:<4} {src}   {tops}")
