"""dev tool (CPU): random block layouts through the kernel bodies under the SIMT emulator (tests/emu), checked against
the plain-C restatement / the system libraries' output.  Long one-byte runs, random bytes, four-letter text and copies
taken from anywhere earlier -- the shapes that found the two bugs fixed at the end of round 2 (a job of the raw / RLE
stage passing for a later block of its frame; an in-ring match wrapping round onto its own source).

usage: python tools/emu_fuzz.py pipeline|corrupt|warp|cta|warp-corrupt|cta-corrupt|encode FIRST_SEED LAST_SEED
  corrupt:  the pipeline over batches in which about half of the frames carry a mutation (flipped bytes, a cut, bytes
            appended): every frame's verdict and, when accepted, bytes are the plain-C restatement's, and the intact
            frames beside them come out right
  pipeline: the phase-split zstd pipeline (early pass, jobs; ZP_EMU_* environment switches of tests/emu apply)
  warp:     the warp-per-block LZ4 and zstd decoders
  cta:      the CTA-per-block LZ4 decoder and the CTA executor of the zstd pipeline (64-thread build;
            EMU_FUZZ_CX_LIB=libcryoemu_cx1024.so: the build with the device's 1 024 threads)
  encode:   the LZ4 and zstd encoders over the same layouts at several sizes; the plain-C restatement decodes them back
  warp-corrupt, cta-corrupt: mutated LZ4 blocks and zstd frames through those decoders, one at a time: the verdict and,
            when accepted, the bytes of the plain-C restatement"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import benchdata                          # noqa: E402
import test_emu_cx as tc                  # noqa: E402
import test_emu_kernels as tk             # noqa: E402
from oracle import port                   # noqa: E402

MiB = 1 << 20
# zstd levels of the frames fed to the decoders: -3 .. 3 (what the GUC's default neighbourhood writes); EMU_FUZZ_LEVELS=lo,hi
# widens it (the GUC allows -5 .. 22: Repeat_Mode tables, single-stream literals, long windows)
LO, HI = (int(x) for x in os.environ.get("EMU_FUZZ_LEVELS", "-3,4").split(","))


def layout(rng, cap, long_runs):
    parts, size = [], 0
    while size < cap:
        kind = int(rng.integers(0, 6))
        if kind == 0:
            p = np.full(int(rng.integers(2_000, 140_000 if long_runs else 40_000)), int(rng.integers(0, 3)) * 7, dtype=np.uint8)
        elif kind == 1:
            p = rng.integers(0, 256, size=int(rng.integers(100, 3000)), dtype=np.uint8)
        elif kind == 2:
            p = rng.integers(97, 101, size=int(rng.integers(500, 9000)), dtype=np.uint8)
        elif kind == 3 and size > 100:
            cur = np.concatenate(parts)
            at = int(rng.integers(0, size - 50))
            p = cur[at: at + int(rng.integers(8, 3000))].copy()
        elif kind == 4:
            p = np.zeros(int(rng.integers(31_000, 34_000) if long_runs else rng.integers(400, 9000)), dtype=np.uint8)
        else:
            p = np.zeros(int(rng.integers(1, 300)), dtype=np.uint8)
        parts.append(p)
        size += p.size
    return np.concatenate(parts)[:cap].copy()


def zstd_compress(buf, level):
    _, zstd = benchdata._libs()
    scratch = np.zeros(buf.size + buf.size // 128 + 4096, dtype=np.uint8)
    got = zstd.ZSTD_compress(scratch.ctypes.data, scratch.size, buf.ctypes.data, buf.size, level)
    assert 0 < got <= scratch.size
    return scratch[:got].copy()


def main():
    what, first, last = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    bad = 0
    for seed in range(first, last):
        rng = np.random.default_rng(99_000 + seed)
        if what == "pipeline":
            n = [5, 8, 9, 14, 16][seed % 5]
            n = int(os.environ.get("EMU_FUZZ_N", n))           # frames per batch
            cap = [MiB, 300 * 1024, 160 * 1024, 640 * 1024][seed % 4]
            plain = [layout(rng, cap, True) for _ in range(n)]
            comp = [zstd_compress(b, int(rng.integers(LO, HI))) for b in plain]
            st, osz, outs, fl = tk._run_pipeline(tk._pipeline_lib(), comp, cap=cap, shift=seed % 16)
            ok = all(st[k] == 0 and osz[k] == cap and np.array_equal(outs[k][:cap], plain[k]) for k in range(n))
            print(seed, "pipeline", n, cap, "ok" if ok else "MISMATCH", "fallback", sum(1 for f in fl if f), flush=True)
        elif what == "corrupt":
            n = [6, 9, 12][seed % 3]
            cap = [160 * 1024, 300 * 1024, 640 * 1024][seed % 3]
            plain = [layout(rng, cap, bool(seed & 1)) for _ in range(n)]
            comp = [zstd_compress(b, int(rng.integers(LO, HI))) for b in plain]
            hit = []
            for k in range(n):
                how = int(rng.integers(0, 8))
                c = comp[k]
                if how == 0:
                    for _ in range(int(rng.integers(1, 4))):
                        c[int(rng.integers(0, c.size))] ^= int(rng.integers(1, 256))
                elif how == 1:
                    c[int(rng.integers(0, min(c.size, 24)))] ^= 1 << int(rng.integers(0, 8))      # headers
                elif how == 2:
                    comp[k] = c[: int(rng.integers(0, c.size))].copy()
                elif how == 3:
                    comp[k] = np.concatenate([c, rng.integers(0, 256, size=int(rng.integers(1, 9)), dtype=np.uint8)])
                hit.append(how < 4)
            want = [port.zstd_decode(c, cap=cap)[:2] for c in comp]
            st, osz, outs, fl = tk._run_pipeline(tk._pipeline_lib(), comp, cap=cap, shift=seed % 16)
            ok = True
            for k in range(n):
                wn, wout = want[k]
                if wn != cap:           # the C-ABI takes blocks of exactly cap bytes: anything else is an error
                    good = st[k] != 0
                else:
                    good = st[k] == 0 and osz[k] == cap and np.array_equal(outs[k][:cap], wout[:cap])
                if not hit[k]:
                    good = good and st[k] == 0 and np.array_equal(outs[k][:cap], plain[k])
                if not good:
                    print("  frame", k, "mutated" if hit[k] else "intact", "status", st[k], "size", osz[k], "port", wn)
                ok = ok and good
            print(seed, "corrupt", n, cap, "ok" if ok else "MISMATCH", "mutated", sum(hit),
                  "accepted", sum(1 for k in range(n) if hit[k] and st[k] == 0), flush=True)
        elif what == "encode":
            L = C.CDLL(os.path.join(ROOT, "tests", "emu", "libcryoemu.so"))
            cap = [MiB, 300_000, 70_001, 4_099, 131_072 + 5, 777_777, 63, 65_536][seed % 8]
            buf = layout(rng, cap, bool(seed & 1))
            ok = True
            for name, fn, arg, dec in (("lz4", L.emu_lz4_encode, 1 + seed % 5, port.lz4_decode),
                                       ("zstd", L.emu_zstd_encode, [1, -5, 3, -1][seed % 4], port.zstd_decode)):
                fn.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_int, C.POINTER(C.c_uint32)]
                fn.restype = C.c_int
                room = cap + cap // 128 + 4096
                out, sz = np.zeros(room, dtype=np.uint8), C.c_uint32(0)
                st = fn(buf.ctypes.data, cap, out.ctypes.data, room, arg, C.byref(sz))
                got, back = dec(out[: sz.value].copy(), cap=cap)[:2]
                good = st == 0 and got == cap and np.array_equal(back[:cap], buf)
                if not good:
                    print("  ", name, "status", st, "size", sz.value, "decoded", got)
                ok = ok and good
            print(seed, "encode", cap, "ok" if ok else "MISMATCH", flush=True)
        else:
            warp = what in ("warp", "warp-corrupt")
            mutate = what.endswith("-corrupt")
            L = C.CDLL(os.path.join(ROOT, "tests", "emu", "libcryoemu.so" if warp else os.environ.get("EMU_FUZZ_CX_LIB", "libcryoemu_cx64.so")))
            cap = ([MiB, 300_000, 777_777, 65_536] if warp else [70_000, 150_000, 40_000, 100_000])[seed % 4]
            buf = layout(rng, cap, warp)
            lz = L.emu_lz4w_decode if warp else L.emu_lz4c_decode
            lz.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint, C.POINTER(C.c_uint32)]
            c = tc.lz4_compress(buf, accel=1 + seed % 4)
            if mutate:
                z = zstd_compress(buf, int(rng.integers(LO, HI)))
                ok = True
                for name, stream, dec, ref in (("lz4", c, lz, port.lz4_decode), ("zstd", z, None, port.zstd_decode)):
                    for _ in range(6):
                        m = stream.copy()
                        how = int(rng.integers(0, 4))
                        if how == 0:
                            for _ in range(int(rng.integers(1, 4))):
                                m[int(rng.integers(0, m.size))] ^= int(rng.integers(1, 256))
                        elif how == 1:
                            m = m[: int(rng.integers(0, m.size))].copy()
                        elif how == 2:
                            m = np.concatenate([m, rng.integers(0, 256, size=int(rng.integers(1, 9)), dtype=np.uint8)])
                        else:
                            m[int(rng.integers(0, min(m.size, 24)))] ^= 1 << int(rng.integers(0, 8))
                        wn, wout = ref(m, cap=cap)[:2]
                        out, sz, fl = np.zeros(cap, dtype=np.uint8), C.c_uint32(0), C.c_uint32(0)
                        src = m.ctypes.data if m.size else None
                        if dec is not None:
                            st = dec(src, m.size, out.ctypes.data, cap, seed % 16, C.byref(sz))
                        elif warp:
                            L.emu_zstdw_decode.argtypes = lz.argtypes
                            st = L.emu_zstdw_decode(src, m.size, out.ctypes.data, cap, seed % 16, C.byref(sz))
                        else:
                            L.emu_zstdc_decode.argtypes = lz.argtypes + [C.POINTER(C.c_uint32)]
                            st = L.emu_zstdc_decode(src, m.size, out.ctypes.data, cap, seed % 16, C.byref(sz), C.byref(fl))
                        good = st != 0 if wn < 0 else \
                            (st == 0 and sz.value == wn and np.array_equal(out[:wn], wout[:wn]))
                        if not good:
                            print("  ", name, "how", how, "status", st, "size", sz.value, "port", wn)
                        ok = ok and good
                print(seed, what, cap, "ok" if ok else "MISMATCH", flush=True)
                bad += not ok
                continue
            out, sz = np.zeros(cap, dtype=np.uint8), C.c_uint32(0)
            st = lz(c.ctypes.data, c.size, out.ctypes.data, cap, seed % 16, C.byref(sz))
            ok = st == 0 and sz.value == cap and np.array_equal(out, buf)
            z = zstd_compress(buf, int(rng.integers(LO, HI)))
            want_n, want = port.zstd_decode(z, cap=cap)[:2]
            assert want_n == cap and np.array_equal(want[:cap], buf)
            out2, sz2, fl = np.zeros(cap, dtype=np.uint8), C.c_uint32(0), C.c_uint32(0)
            if warp:
                L.emu_zstdw_decode.argtypes = lz.argtypes
                st2 = L.emu_zstdw_decode(z.ctypes.data, z.size, out2.ctypes.data, cap, seed % 16, C.byref(sz2))
            else:
                L.emu_zstdc_decode.argtypes = lz.argtypes + [C.POINTER(C.c_uint32)]
                st2 = L.emu_zstdc_decode(z.ctypes.data, z.size, out2.ctypes.data, cap, seed % 16, C.byref(sz2), C.byref(fl))
            ok = ok and st2 == 0 and sz2.value == cap and np.array_equal(out2, buf)
            print(seed, what, cap, "ok" if ok else "MISMATCH", flush=True)
        bad += not ok
    print("mismatches:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
