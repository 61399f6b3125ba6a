/*
 * tests/emu/emu_cx.cpp -- TEST INFRASTRUCTURE ONLY.
 * The CTA-per-block decoders (cryo_cx.cuh, lz4_decode_c.cuh, zstd_decode_c.cuh) on the CPU through
 * cuda_emu.h.  Built twice: with a small CTA (-DCX_THREADS=64: short parse regions, many rounds and
 * chunk cuts per block) and with the product's 1024 threads.
 */
#define CRYO_EMU 1
#include "cuda_emu.h"
#include "../../pg_cryogen_b200/csrc/lz4_decode_c.cuh"
#include "../../pg_cryogen_b200/csrc/zstd_decode_c.cuh"

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

/* one zstd frame through stages 1-3 of the pipeline and the CTA-per-frame stage 4; a flagged frame goes to
 * the warp-per-frame decoder like in the library.  *flag_out: the pipeline declined the frame. */
extern "C" int
emu_zstdc_decode(const uint8_t *src, uint32_t csize, uint8_t *dst, uint32_t cap, unsigned shift, uint32_t *out_size,
                 uint32_t *flag_out)
{
    static uint32_t predef[ZSW_PREDEF_CELLS];
    static bool have_predef = false;
    const uint32_t stride = (cap + 15u) & ~15u;
    std::vector<uint8_t> ibuf((size_t) csize + 512, 0xEE), obuf((size_t) stride + 512, 0xAA), scr(ZSTDD_SCRATCH_BYTES + 128, 0x77);
    uint8_t *ib = (uint8_t *) ((((uintptr_t) ibuf.data() + 63) & ~(uintptr_t) 63) + 64);
    uint8_t *o = (uint8_t *) ((((uintptr_t) obuf.data() + 63) & ~(uintptr_t) 63) + 64);
    uint8_t *sc = (uint8_t *) ((((uintptr_t) scr.data() + 63) & ~(uintptr_t) 63));
    int32_t status = -1, method = ZP_METHOD_ZSTD;
    uint64_t off = shift & 15;

    if (!have_predef)
    {
        emu::launch(dim3(1), dim3(32), 2048, [&]() { zsw_build_predef(predef, CRYO_SMEM_BASE(), threadIdx.x); });
        have_predef = true;
    }
    if (csize)
        memcpy(ib + off, src, csize);
    const uint64_t lit_stride = zp_lit_stride(cap), seq_cap = zp_seq_cap(1, cap);
    std::vector<uint32_t> fr(ZP_FF, 0xCDCDCDCD), blk((size_t) ZP_MAXB * ZP_BF, 0xCDCDCDCD), flag(1, 0xCDCDCDCD), pf_done(1, 0), cxlist(1, 0);
    std::vector<uint64_t> seqbase(1, 0), seq(seq_cap + 8, 0x7777777777777777ull);
    std::vector<uint8_t>  lit(lit_stride + 64, 0x99);
    std::vector<uint16_t> huftab((size_t) ZP_MAXB * 2048, 0x3333);
    std::vector<uint32_t> fsetab((size_t) ZP_MAXB * ZP3_CELLS, 0x44444444);
    unsigned long long    seq_alloc = 0;
    uint32_t              cxcount = 0;
    ZpArgs a;

    a.methods = &method;
    a.src = ib;
    a.src_off = &off;
    a.src_size = &csize;
    a.dst = o;
    a.dst_stride = stride;
    a.cap = cap;
    a.n = 1;
    a.out_size = out_size;
    a.status = &status;
    a.fr = fr.data();
    a.blk = blk.data();
    a.flag = flag.data();
    a.seqbase = seqbase.data();
    a.seq_alloc = &seq_alloc;
    a.cxcount = &cxcount;
    a.cxlist = cxlist.data();
    a.pf_done = pf_done.data();
    a.pf_inflight = 2;
    a.jobs = nullptr;
    a.pf_expect = nullptr;
    a.exec_warps = 0;
    a.early_frames = 0;
    a.pf_hint = 4u;                     /* every frame to the CTA-per-frame stage 4 */
    a.lit = (uint8_t *) ((((uintptr_t) lit.data() + 15) & ~(uintptr_t) 15));
    a.lit_stride = lit_stride;
    a.seq = seq.data();
    a.seq_cap = seq_cap;
    a.predef = predef;
    a.huftab = huftab.data();
    a.fsetab = fsetab.data();
    const unsigned split = 2;           /* warps per group of the lane-serial stages (ZP_FOR_GROUP_BLOCKS) */

    emu::launch(dim3(1), dim3(32), 0, [&]() {
        if (threadIdx.x == 0)
            zp_stage1(a, 0);
    });
    emu::launch(dim3(split), dim3(32 * ZP2A_WARPS), ZP2A_SMEM, [&]() {
        ZP_FOR_GROUP_BLOCKS_CTA(a, blockIdx.x, split, ZPF_HUFMASK, threadIdx.x, zp_stage2a(a, g, j, CRYO_SMEM_BASE(), threadIdx.x));
    });
    emu::launch(dim3(split), dim3(32), ZP2B_SMEM, [&]() {
        ZP_FOR_GROUP_BLOCKS(a, blockIdx.x, split, ZPF_HUFMASK, threadIdx.x, zp_stage2b(a, g, j, CRYO_SMEM_BASE(), threadIdx.x));
    });
    emu::launch(dim3(split), dim3(32 * ZP3A_WARPS), ZP3A_SMEM, [&]() {
        ZP_FOR_GROUP_BLOCKS_CTA(a, blockIdx.x, split, ZPF_SEQMASK, threadIdx.x, zp_stage3a(a, g, j, CRYO_SMEM_BASE(), threadIdx.x));
    });
    emu::launch(dim3(split), dim3(32), ZP3B_SMEM(ZP3B_SMALL, ZP3B_SMALL_LANES), [&]() {
        ZP_FOR_GROUP_BLOCKS(a, blockIdx.x, split, ZPF_SEQMASK, threadIdx.x,
                            (zp_stage3b<ZP3B_SMALL, 0, ZP3B_SMALL_LANES>(a, g, j, CRYO_SMEM_BASE(), threadIdx.x)));
    });
    emu::launch(dim3(split), dim3(32), ZP3B_SMEM(ZP3B_LARGE, ZP_G), [&]() {
        ZP_FOR_GROUP_BLOCKS(a, blockIdx.x, split, ZPF_SEQMASK, threadIdx.x,
                            (zp_stage3b<ZP3B_LARGE, ZP3B_SMALL, ZP_G>(a, g, j, CRYO_SMEM_BASE(), threadIdx.x)));
    });
    if (getenv("CRYO_EMU_TRACE"))
        fprintf(stderr, "after stages 1-3: flag %u nblk %u route %u\n", flag[0], fr[0], fr[3]);
    if (flag[0] == 0 && fr[0] != 0 && fr[3] == 1)
        emu::launch(dim3(1), dim3(CX_THREADS), ZC_SMEM, [&]() { zp_stage4_cx(a, 0, CRYO_SMEM_BASE(), threadIdx.x); });
    *flag_out = flag[0];
    if (flag[0])
        emu::launch(dim3(1), dim3(32), ZSW_PER_WARP, [&]() {
            zstdw_decode_frame(ib + off, csize, o, cap, out_size, &status, sc, predef, CRYO_SMEM_BASE(), threadIdx.x);
        });
    memcpy(dst, o, cap);
    for (int i = 1; i <= 64; i++)
        if (o[-i] != 0xAA || o[stride + i - 1] != 0xAA)
            return -100;
    return status;
}
