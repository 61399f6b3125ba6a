"""dev probe: phase cycle counters of the CTA-per-block LZ4 decoder (needs a -DCX_PROF build of libcryogpu.so:
   nvcc ... -DCX_PROF -o gpurun_out/libcryogpu_prof.so; run with CRYOGPU_LIB=that path)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from pg_cryogen_b200 import codec
if os.environ.get("CRYOGPU_LIB"):
    codec.lib_path = lambda: os.environ["CRYOGPU_LIB"]
from pg_cryogen_b200 import CryoGPU, blockgen as bg
from gpu_util import decode_device
from oracle import ref
NAMES = {0: "stage", 1: "walk", 2: "link", 3: "mark", 4: "emit", 5: "load+bar", 6: "scan", 7: "pre", 8: "cut+checks", 9: "literals",
         10: "depsearch", 11: "matches", 12: "drain", 13: "bulk", 14: "finish", 16: "#rounds", 17: "#chunks", 18: "#bulk", 19: "#markpasses", 20: "d.pend", 21: "d.lds", 22: "d.stg"}
g = CryoGPU(0)
L = g.lib
for method, kind, pl in ((0, "S", "hex"), (0, "M", "hex"), (0, "D", "hex"), (0, "D", "lowcard")) + ((1, "D", "lowcard"), (1, "M", "hex")) * int(os.environ.get("CXPROF_ZSTD", "0")):
    blk = bg.make_block(kind, pl, 3)[None]
    c = ref.compress(method, 1, blk)[0]
    decode_device(g, method, c)
    L.cryogpu_debug_cxprof(None, 1)
    out, osz, st = decode_device(g, method, c)
    buf = (C.c_ulonglong * 32)()
    L.cryogpu_debug_cxprof(buf, 0)
    tot = sum(buf[i] for i in range(16)) + sum(buf[i] for i in range(20, 24))
    print(f"method {method} {kind}/{pl}: ok={st[0] == 0 and np.array_equal(out[0], blk[0])} total {tot} cycles = {tot / 1.9e3:.0f} us")
    print("   " + "  ".join(f"{NAMES.get(i, i)}={buf[i]}" for i in range(24) if buf[i]))
