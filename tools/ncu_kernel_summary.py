"""Print one line per kernel of an .ncu-rep: duration, DRAM bytes, instructions, issue-active, occupancy, launch shape.
usage: python tools/ncu_kernel_summary.py report.ncu-rep [label]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; label = sys.argv[2] if len(sys.argv) > 2 else rep
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
col = lambda n: hdr.index(n)
tb = lambda v, u: float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
for r in rows[2:]:
    g = lambda n: (r[col(n)], units[col(n)])
    t = float(g("gpu__time_duration.sum")[0]) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}[g("gpu__time_duration.sum")[1]]
    rd = tb(*g("dram__bytes_read.sum")); wr = tb(*g("dram__bytes_write.sum"))
    print(f"{label}: {r[col('Kernel Name')].split('(')[0]:18s} {t:9.1f} us  dram rd {rd/1e6:8.1f} MB wr {wr/1e6:8.1f} MB "
          f"({(rd+wr)/t/1e3:6.1f} GB/s)  inst {float(g('smsp__inst_executed.sum')[0])/1e6:8.1f} M  "
          f"issue {float(g('smsp__issue_active.avg.pct_of_peak_sustained_active')[0]):5.1f}%  "
          f"warps {float(g('sm__warps_active.avg.pct_of_peak_sustained_active')[0]):5.1f}%  regs {g('launch__registers_per_thread')[0]}  "
          f"grid {g('launch__grid_size')[0]} x {g('launch__block_size')[0]}  smem {float(g('launch__shared_mem_per_block_dynamic')[0]):.1f} KB")
