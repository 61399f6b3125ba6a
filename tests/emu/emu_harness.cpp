/*
 * tests/emu/emu_harness.cpp -- TEST INFRASTRUCTURE ONLY.
 * Runs the kernel bodies from pg_cryogen_b200/csrc/*.cuh on the CPU through
 * cuda_emu.h so that the `-m "not gpu"` suite can check kernel logic bit-exactly
 * against the oracle without a GPU.  Never part of libcryogpu.so.
 */
#define CRYO_EMU 1
#include "cuda_emu.h"
#include "../../pg_cryogen_b200/csrc/lz4_decode.cuh"
#include "../../pg_cryogen_b200/csrc/zstd_decode.cuh"

#include <vector>

namespace {
/* 16-byte aligned copy with `shift` bytes of misalignment and 64 bytes of slack */
struct Padded
{
    std::vector<uint8_t> buf;
    uint8_t *p;
    Padded(const uint8_t *src, size_t n, unsigned shift, uint8_t fill = 0xEE)
    {
        buf.assign(n + 256, fill);
        uintptr_t a = ((uintptr_t) buf.data() + 63) & ~(uintptr_t) 63;
        p = (uint8_t *) a + 64 + (shift & 15);
        if (n)
            memcpy(p, src, n);
    }
};
}

extern "C" int
emu_lz4_decode(const uint8_t *src, uint32_t csize, uint8_t *dst, uint32_t cap, unsigned shift,
               uint32_t *out_size)
{
    Padded  in(src, csize, shift);
    Padded  out(nullptr, 0, 0);
    int32_t status = -1;

    out.buf.assign(cap + 256, 0xAA);
    uint8_t *o = (uint8_t *) ((((uintptr_t) out.buf.data() + 63) & ~(uintptr_t) 63) + 64);
    emu::launch(dim3(1), dim3(LZ4D_THREADS), LZ4D_SMEM, [&]() {
        lz4_decode_block(in.p, csize, o, cap, out_size, &status);
    });
    memcpy(dst, o, cap);
    /* the kernel must not write outside [o, o + cap) */
    for (int i = 1; i <= 64; i++)
        if (o[-i] != 0xAA || o[cap + i - 1] != 0xAA)
            return -100;
    return status;
}

extern "C" int
emu_zstd_decode(const uint8_t *src, uint32_t csize, uint8_t *dst, uint32_t cap, unsigned shift,
                uint32_t *out_size)
{
    Padded  in(src, csize, shift);
    std::vector<uint8_t> obuf(cap + 256, 0xAA), scr(ZSTDD_SCRATCH_BYTES + 128, 0x77);
    int32_t status = -1;
    uint8_t *o = (uint8_t *) ((((uintptr_t) obuf.data() + 63) & ~(uintptr_t) 63) + 64);
    uint8_t *sc = (uint8_t *) ((((uintptr_t) scr.data() + 63) & ~(uintptr_t) 63));

    emu::launch(dim3(1), dim3(ZSTDD_THREADS), ZSTDD_SMEM, [&]() {
        zstd_decode_frame(in.p, csize, o, cap, out_size, &status, sc);
    });
    memcpy(dst, o, cap);
    for (int i = 1; i <= 64; i++)
        if (o[-i] != 0xAA || o[cap + i - 1] != 0xAA)
            return -100;
    return status;
}
