/*
 * tests/emu/emu_pages.cpp -- TEST INFRASTRUCTURE ONLY.
 * The page-chain and tuple-walk kernel bodies (pg_cryogen_b200/csrc/cryo_pages.cuh) on the CPU through cuda_emu.h,
 * launched as cryogpu.cu launches them (one CTA of 256 threads per cryo block; one warp per decoded block), so the
 * `-m "not gpu"` suite can hold them against oracle/cryo_pages.c.  Never part of libcryogpu.so.
 */
#define CRYO_EMU 1
#include "cuda_emu.h"
#include "../../pg_cryogen_b200/csrc/cryo_pages.cuh"

#include <vector>

/* k_pages_gather: n cryo blocks; chain entries [chain_off[b], chain_off[b + 1]) of slot / blkno; comp holds PG_PAGE bytes
 * per chain entry and is 16-byte aligned, as are the pages */
extern "C" int
emu_pages_gather(const uint8_t *pages, const uint32_t *slot, const uint32_t *blkno, const uint32_t *chain_off, uint32_t n,
                 uint8_t *comp, uint64_t *src_off, uint32_t *src_size, int32_t *dec_method, int32_t *hdr_method,
                 int32_t *chain_status, uint32_t max_csize)
{
    if (((uintptr_t) pages | (uintptr_t) comp) & 15u)
        return -1;
    emu::launch(dim3(n), dim3(256), 0, [&]() {
        const uint32_t b = blockIdx.x;

        pg_gather_block(pages, slot, blkno, chain_off[b], chain_off[b + 1], comp, src_off + b, src_size + b, dec_method + b,
                        hdr_method + b, chain_status + b, max_csize, threadIdx.x, 256);
    });
    return 0;
}

/* k_pages_split without the encoder's status: npages[b] = 0 when the pages do not fit cap_pages */
extern "C" int
emu_pages_split(const uint8_t *comp, uint64_t comp_stride, const uint32_t *comp_size, uint32_t n, uint32_t method,
                uint32_t xid, const uint32_t *blkno, uint32_t cap_pages, uint8_t *pages, uint64_t pages_stride,
                uint32_t *npages)
{
    if (((uintptr_t) pages | (uintptr_t) comp | comp_stride | pages_stride) & 15u)
        return -1;
    emu::launch(dim3(n), dim3(256), 0, [&]() {
        const uint32_t b = blockIdx.x;
        const uint32_t np = pg_split_block(comp + b * comp_stride, comp_size[b], method, xid, blkno + (size_t) b * cap_pages,
                                           cap_pages, pages + b * pages_stride, threadIdx.x, 256);

        if (threadIdx.x == 0)
            npages[b] = np;
    });
    return 0;
}

/* k_tuple_stats: one warp per block, 8 blocks per CTA */
extern "C" int
emu_tuple_stats(const uint8_t *blocks, uint64_t stride, uint32_t block_size, uint32_t n, uint32_t *ntuples,
                unsigned long long *tuple_bytes, int32_t *valid)
{
    emu::launch(dim3((n + 7u) / 8u), dim3(256), 0, [&]() {
        const uint32_t b = blockIdx.x * 8u + (threadIdx.x >> 5), lane = threadIdx.x & 31u;

        if (b < n)
            pg_tuple_stats(blocks + b * stride, block_size, ntuples + b, tuple_bytes + b, valid + b, lane);
    });
    return 0;
}
