"""dev: run the zstd encoder kernel body under the CPU emulator and round-trip the frame through
the reference's libzstd (oracle/_ref).  usage: emu_zstd_enc.py [level ...]"""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import ref
from pg_cryogen_b200 import blockgen as bg

L = C.CDLL(os.path.join(os.path.dirname(__file__), "..", "tests", "emu", "libcryoemu.so"))
L.emu_zstd_encode.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_int, C.POINTER(C.c_uint32)]
L.emu_zstd_encode.restype = C.c_int
MiB = 1 << 20

def enc(blk, level):
    cap = ref.compress_bound(1)
    out = np.zeros(cap, dtype=np.uint8); sz = C.c_uint32(0)
    st = L.emu_zstd_encode(blk.ctypes.data, blk.size, out.ctypes.data, cap, level, C.byref(sz))
    return st, out[: sz.value].copy()

if __name__ == "__main__":
    levels = [int(a) for a in sys.argv[1:]] or [1]
    cases = [tuple(c.split("/")) for c in os.environ["CASES"].split(",")] if os.environ.get("CASES") else [("S", "hex"), ("S", "lowcard"), ("M", "hex"), ("M", "lowcard"), ("D", "hex"), ("D", "lowcard"), ("D", "random")]
    for lv in levels:
        for kind, pl in cases:
            blk = bg.make_block(kind, pl, 11)
            t0 = time.time(); st, c = enc(blk, lv); dt = time.time() - t0
            back, ok = ref.decompress_one(1, c) if st == 0 else (None, False)
            good = ok and np.array_equal(back, blk)
            rsz = int(ref.compress(1, lv, blk)[1][0])
            print(os.environ.get("ZSE_DBG_MM","-"), os.environ.get("ZSE_DBG_STEP","-"), end=" "); print(f"level {lv:3d} {kind}/{pl:8s} status={st} size={len(c):8d} ref={rsz:8d} ratio={len(c)/rsz:6.3f} roundtrip={'OK' if good else 'FAIL'}  ({dt:.1f}s)", flush=True)
