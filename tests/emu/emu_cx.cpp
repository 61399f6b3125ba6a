/*
 * tests/emu/emu_cx.cpp -- TEST INFRASTRUCTURE ONLY.
 * The CTA-per-block decoders (cryo_cx.cuh, lz4_decode_c.cuh, zstd_decode_c.cuh) on the CPU through
 * cuda_emu.h.  Built twice: with a small CTA (-DCX_THREADS=64: short parse regions, many rounds and
 * chunk cuts per block) and with the product's 1024 threads.
 */
#define CRYO_EMU 1
#include "cuda_emu.h"
#include "../../pg_cryogen_b200/csrc/lz4_decode_c.cuh"

#include <vector>

extern "C" int
emu_cx_threads(void)
{
    return (int) CX_THREADS;
}

extern "C" int
emu_lz4c_decode(const uint8_t *src, uint32_t csize, uint8_t *dst, uint32_t cap, unsigned shift, uint32_t *out_size)
{
    std::vector<uint8_t> ibuf((size_t) csize + 512, 0xEE), obuf((size_t) cap + 512, 0xAA);
    std::vector<unsigned long long> gseq(LZ4C_SEQCAP + 16);
    uint8_t *ip = (uint8_t *) ((((uintptr_t) ibuf.data() + 63) & ~(uintptr_t) 63) + 64 + (shift & 15));
    uint8_t *o = (uint8_t *) ((((uintptr_t) obuf.data() + 63) & ~(uintptr_t) 63) + 64);
    int32_t status = -1;

    if (csize)
        memcpy(ip, src, csize);
    emu::launch(dim3(1), dim3(CX_THREADS), LZ4C_SMEM, [&]() {
        lz4c_decode_block(ip, csize, o, cap, out_size, &status, CRYO_SMEM_BASE(), gseq.data(), threadIdx.x);
    });
    memcpy(dst, o, cap);
    for (int i = 1; i <= 64; i++)
        if (o[-i] != 0xAA || o[((cap + 15u) & ~15u) + i - 1] != 0xAA)
            return -100;                /* wrote outside the block */
    return status;
}
