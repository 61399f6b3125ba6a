/*
 * tests/emu/emu_harness.cpp -- TEST INFRASTRUCTURE ONLY.
 * Runs the kernel bodies from pg_cryogen_b200/csrc/*.cuh on the CPU through
 * cuda_emu.h so that the `-m "not gpu"` suite can check kernel logic bit-exactly
 * against the oracle without a GPU.  Never part of libcryogpu.so.
 */
#define CRYO_EMU 1
#include "cuda_emu.h"
#include "../../pg_cryogen_b200/csrc/lz4_decode_w.cuh"
#include "../../pg_cryogen_b200/csrc/zstd_decode_w.cuh"
#include "../../pg_cryogen_b200/csrc/zstd_decode_p.cuh"
#include "../../pg_cryogen_b200/csrc/lz4_encode.cuh"
#include "../../pg_cryogen_b200/csrc/zstd_encode.cuh"

#include <vector>
#include <stdio.h>
#include <stdlib.h>

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
emu_lz4w_decode(const uint8_t *src, uint32_t csize, uint8_t *dst, uint32_t cap, unsigned shift,
                uint32_t *out_size)
{
    Padded  in(src, csize, shift);
    std::vector<uint8_t> obuf(2 * (size_t) cap + 512, 0xAA);
    int32_t status[2] = {-1, -1};
    uint32_t osz[2] = {0, 0};
    uint8_t *o = (uint8_t *) ((((uintptr_t) obuf.data() + 63) & ~(uintptr_t) 63) + 64);
    uint32_t stride = (cap + 15u) & ~15u;

    emu::launch(dim3(1), dim3(LZ4W_THREADS), LZ4W_SMEM, [&]() {
        uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (warp < 2)
            lz4w_decode_block(in.p, csize, o + warp * (size_t) stride, cap, osz + warp, status + warp,
                              CRYO_SMEM_BASE() + warp * LZ4W_PER_WARP, lane);
    });
    if (status[0] != status[1] || osz[0] != osz[1] || memcmp(o, o + stride, cap) != 0)
        return -101;
    memcpy(dst, o, cap);
    *out_size = osz[0];
    for (int i = 1; i <= 64; i++)
        if (o[-i] != 0xAA)
            return -100;
    return status[0];
}

/* warp-per-frame zstd decoder; the predefined tables are built by the same device code */
extern "C" int
emu_zstdw_decode(const uint8_t *src, uint32_t csize, uint8_t *dst, uint32_t cap, unsigned shift,
                 uint32_t *out_size)
{
    Padded  in(src, csize, shift);
    std::vector<uint8_t> obuf(2 * (size_t) cap + 512, 0xAA), scr(2 * ZSTDD_SCRATCH_BYTES + 128, 0x77);
    static uint32_t predef[ZSW_PREDEF_CELLS];
    static bool have_predef = false;
    int32_t status[2] = {-1, -1};
    uint32_t osz[2] = {0, 0};
    uint8_t *o = (uint8_t *) ((((uintptr_t) obuf.data() + 63) & ~(uintptr_t) 63) + 64);
    uint8_t *sc = (uint8_t *) ((((uintptr_t) scr.data() + 63) & ~(uintptr_t) 63));
    uint32_t stride = (cap + 15u) & ~15u;

    if (!have_predef)
    {
        emu::launch(dim3(1), dim3(32), 2048, [&]() { zsw_build_predef(predef, CRYO_SMEM_BASE(), threadIdx.x); });
        have_predef = true;
    }
    emu::launch(dim3(1), dim3(ZSW_THREADS), ZSW_SMEM, [&]() {
        uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (warp < 2)
            zstdw_decode_frame(in.p, csize, o + warp * (size_t) stride, cap, osz + warp, status + warp,
                               sc + warp * ZSTDD_SCRATCH_BYTES, predef,
                               CRYO_SMEM_BASE() + warp * ZSW_PER_WARP, lane);
    });
    if (status[0] != status[1] || osz[0] != osz[1] || memcmp(o, o + stride, cap) != 0)
        return -101;
    memcpy(dst, o, cap);
    *out_size = osz[0];
    for (int i = 1; i <= 64; i++)
        if (o[-i] != 0xAA)
            return -100;
    return status[0];
}

/* serial vs warp-parallel FSE table build on the same counts (returns number of differing cells) */
extern "C" int
emu_fse_build_compare(const int16_t *counts, int nsym, int log)
{
    static uint32_t a[512], b[512];
    static uint16_t nexts[64];
    int diff = 0;

    fse_build_table(a, counts, nsym, log, nexts);
    emu::launch(dim3(1), dim3(32), 4096, [&]() {
        uint8_t *sm = CRYO_SMEM_BASE();
        int16_t *c = reinterpret_cast<int16_t *>(sm);
        for (int i = threadIdx.x; i < nsym; i += 32)
            c[i] = counts[i];
        __syncwarp();
        fse_build_table_warp(reinterpret_cast<uint32_t *>(sm + 1024), c, nsym, log,
                             reinterpret_cast<uint16_t *>(sm + 256),
                             reinterpret_cast<uint16_t *>(sm + 512), threadIdx.x);
        __syncwarp();
        for (int i = threadIdx.x; i < (1 << log); i += 32)
            b[i] = reinterpret_cast<uint32_t *>(sm + 1024)[i];
    });
    for (int i = 0; i < (1 << log); i++)
        diff += a[i] != b[i];
    return diff;
}

extern "C" int
emu_lz4_encode(const uint8_t *src, uint32_t n, uint8_t *dst, uint32_t dst_cap, int accel,
               uint32_t *dst_size)
{
    std::vector<uint8_t> ibuf((size_t) n + 256, 0xEE), obuf((size_t) dst_cap + 256, 0xAA),
                         scr(lz4e_scratch_bytes(n) + 256, 0x55);
    uint8_t *in = (uint8_t *) ((((uintptr_t) ibuf.data() + 63) & ~(uintptr_t) 63) + 64);
    uint8_t *o = (uint8_t *) ((((uintptr_t) obuf.data() + 63) & ~(uintptr_t) 63) + 64);
    uint8_t *sc = (uint8_t *) ((((uintptr_t) scr.data() + 63) & ~(uintptr_t) 63));
    int32_t status = -1;

    uint32_t queue[65] = {0};           /* [0]: next work item; [64]: finished segments of the block */

    memcpy(in, src, n);
    emu::launch(dim3(1), dim3(LZ4E_THREADS), LZ4E_SMEM, [&]() {
        lz4_encode_worker(in, 0, n, o, 0, dst_cap, accel, dst_size, &status, sc, 0, 1, queue, queue + 64,
                          reinterpret_cast<uint16_t *>(CRYO_SMEM_BASE() + (threadIdx.x >> 5) * LZ4E_HASH_BYTES), threadIdx.x & 31);
    });
    if (status == 0)
        memcpy(dst, o, *dst_size);
    for (int i = 1; i <= 64; i++)
        if (o[-i] != 0xAA || o[dst_cap + i - 1] != 0xAA)
            return -100;
    return status;
}

extern "C" int
emu_zstd_encode(const uint8_t *src, uint32_t n, uint8_t *dst, uint32_t dst_cap, int level,
                uint32_t *dst_size)
{
    std::vector<uint8_t> ibuf((size_t) n + 256, 0xEE), obuf((size_t) dst_cap + 256, 0xAA),
                         scr(zstde_scratch_bytes(n) + 256, 0x55);
    uint8_t *in = (uint8_t *) ((((uintptr_t) ibuf.data() + 63) & ~(uintptr_t) 63) + 64);
    uint8_t *o = (uint8_t *) ((((uintptr_t) obuf.data() + 63) & ~(uintptr_t) 63) + 64);
    uint8_t *sc = (uint8_t *) ((((uintptr_t) scr.data() + 63) & ~(uintptr_t) 63));
    int32_t status = -1;

    uint32_t queue[65] = {0};           /* [0]: next work item; [64]: finished blocks of the frame */
    std::vector<uint32_t> bmeta(ZSE_MAXBLK, 0xCDCDCDCD);

    memcpy(in, src, n);
    emu::launch(dim3(1), dim3(ZSTDE_THREADS), ZSTDE_SMEM, [&]() {
        const uint32_t warp = threadIdx.x >> 5;

        zstd_encode_worker(in, 0, n, o, 0, dst_cap, level, dst_size, &status, sc + (size_t) warp * ZSE_SCR_PER_WARP, 1, queue,
                           queue + 64, bmeta.data(), CRYO_SMEM_BASE() + warp * ZSE_PER_WARP, threadIdx.x & 31);
    });
    if (status == 0)
        memcpy(dst, o, *dst_size);
    for (int i = 1; i <= 64; i++)
        if (o[-i] != 0xAA || o[dst_cap + i - 1] != 0xAA)
            return -100;
    return status;
}

/*
 * group decoder (8 lanes per frame): up to ZSG_GROUPS different streams decoded side by side
 * in one CTA, so the groups of a warp diverge as they do on mixed batches.
 * srcs/csizes/dsts/out_sizes/statuses are arrays of n entries (n <= ZSG_GROUPS).
 */
extern "C" int
emu_zstdp_decode_multi(int n, const uint8_t *const *srcs, const uint32_t *csizes, uint8_t *const *dsts,
                       uint32_t cap, unsigned shift, uint32_t *out_sizes, int32_t *statuses,
                       uint32_t *flags)
{
    static uint32_t predef[ZSW_PREDEF_CELLS];
    static bool have_predef = false;
    const uint32_t stride = (cap + 15u) & ~15u;
    std::vector<uint8_t> obuf((size_t) n * stride + 512, 0xAA), scr(ZSTDD_SCRATCH_BYTES + 128, 0x77);
    uint8_t *o = (uint8_t *) ((((uintptr_t) obuf.data() + 63) & ~(uintptr_t) 63) + 64);
    uint8_t *sc = (uint8_t *) ((((uintptr_t) scr.data() + 63) & ~(uintptr_t) 63));

    if (n < 1)
        return -102;
    if (!have_predef)
    {
        emu::launch(dim3(1), dim3(32), 2048, [&]() { zsw_build_predef(predef, CRYO_SMEM_BASE(), threadIdx.x); });
        have_predef = true;
    }
    /* pack the inputs the way the library does: one buffer, per-frame offsets, odd alignments */
    std::vector<uint64_t> off(n);
    std::vector<int32_t>  methods(n, ZP_METHOD_ZSTD);
    size_t total = 64;

    for (int i = 0; i < n; i++)
    {
        off[i] = total + ((shift + 3 * i) & 15);
        total = (off[i] + csizes[i] + 15 + 16) & ~(size_t) 15;
    }
    std::vector<uint8_t> inbuf(total + 256, 0xEE);
    uint8_t *ib = (uint8_t *) ((((uintptr_t) inbuf.data() + 63) & ~(uintptr_t) 63));

    for (int i = 0; i < n; i++)
        if (csizes[i])
            memcpy(ib + off[i], srcs[i], csizes[i]);
    const uint64_t lit_stride = zp_lit_stride(cap), seq_cap = zp_seq_cap((uint64_t) n, cap);
    std::vector<uint32_t> cxlist((size_t) n, 0);
    std::vector<uint32_t> fr((size_t) n * ZP_FF, 0xCDCDCDCD), blk((size_t) n * ZP_MAXB * ZP_BF, 0xCDCDCDCD),
                          flag(n, 0xCDCDCDCD);
    std::vector<uint64_t> seqbase(n, 0), seq(seq_cap + 8, 0x7777777777777777ull);
    std::vector<uint8_t>  lit((size_t) n * lit_stride + 64, 0x99);
    std::vector<uint16_t> huftab((size_t) n * ZP_MAXB * 2048, 0x3333);
    std::vector<uint32_t> fsetab((size_t) n * ZP_MAXB * ZP3_CELLS, 0x44444444);
    unsigned long long    ctl[8] = {0, 0, 0, 0, 0, 0, 0, 0};    /* seq_alloc and the control words behind it */
    std::vector<uint32_t> jobs((size_t) n * ZP_JOBS * 4, 0), pf_expect(n, 0xCDCDCDCD);
    uint32_t              cxcount = 0;
    std::vector<uint32_t> pf_done(n, 0xCDCDCDCD);
    ZpArgs a;

    a.methods = methods.data();
    a.src = ib;
    a.src_off = off.data();
    a.src_size = csizes;
    a.dst = o;
    a.dst_stride = stride;
    a.cap = cap;
    a.n = (uint32_t) n;
    a.out_size = out_sizes;
    a.status = statuses;
    a.fr = fr.data();
    a.blk = blk.data();
    a.flag = flag.data();
    a.seqbase = seqbase.data();
    a.seq_alloc = ctl;
    a.jobs = getenv("ZP_EMU_NO_JOBS") ? nullptr : jobs.data();
    a.pf_expect = a.jobs ? pf_expect.data() : nullptr;
    a.exec_warps = (((unsigned) n + ZP4_WARPS - 1) / ZP4_WARPS) * ZP4_WARPS;
    a.cxcount = &cxcount;
    a.cxlist = cxlist.data();
    a.pf_done = pf_done.data();
    /* every frame to the warp stage 4 (the CTA stage: emu_cx.cpp); bit 4: the early pass of stage 0 runs (ZP_EMU_NO_EARLY: not) */
    const bool early_pass = getenv("ZP_EMU_NO_EARLY") == nullptr;
    a.pf_hint = 8u;
    a.pf_inflight = 2;
    a.early_frames = early_pass ? ((uint32_t) n + 1u) / 2u : 0u;     /* the last half of the batch */
    a.lit = (uint8_t *) ((((uintptr_t) lit.data() + 15) & ~(uintptr_t) 15));
    a.lit_stride = lit_stride;
    a.seq = seq.data();
    a.seq_cap = seq_cap;
    a.predef = predef;
    a.huftab = huftab.data();
    a.fsetab = fsetab.data();

    const unsigned ngroups = ((unsigned) n + ZP_G - 1) / ZP_G;
    /* warps per group of the lane-serial stages (ZP_FOR_GROUP_BLOCKS): 1, 3 or 16 by the batch size, so the tests cover each */
    const unsigned split = n % 3 == 0 ? 1u : n % 3 == 1 ? 3u : ZP_MAXB;

    emu::launch(dim3(((unsigned) n + 31) / 32), dim3(32), 0, [&]() {
        const uint32_t f = blockIdx.x * 32 + threadIdx.x;
        if (f < a.n)
            zp_stage1(a, f);
    });
    /* ZP_EMU_LATE_PREFILL: stage 0 runs after stage 4 (as if it were scheduled late), so stage 4
     * times out waiting for it and writes the blocks it needs itself */
    const bool late_prefill = getenv("ZP_EMU_LATE_PREFILL") != nullptr;
    auto prefill = [&]() {
        emu::launch(dim3(3), dim3(128), 64, [&]() {
            uint32_t *pos = reinterpret_cast<uint32_t *>(CRYO_SMEM_BASE());

            for (uint32_t f = blockIdx.x; f < a.n; f += 3)
            {
                __syncthreads();
                if (threadIdx.x < 32)
                {
                    const uint32_t at = zp_frame_positions_warp<false>(a, f, threadIdx.x);

                    if (threadIdx.x < ZP_MAXB)
                        pos[threadIdx.x] = at;
                }
                __syncthreads();
                if (threadIdx.x == 0)
                {
                    uint32_t serial[ZP_MAXB];

                    zp_frame_positions_late(a, f, serial);
                    for (uint32_t j = 0; j < ZP_MAXB; j++)
                        if (serial[j] != pos[j])
                            abort();            /* the warp version and the serial one must agree */
                }
                __syncthreads();
                for (uint32_t j = 0; j < fr[(size_t) f * ZP_FF]; j++)
                {
                    if (pos[j] == ~0u)
                        continue;
                    zp_stage0(a, (f << 8) | j, pos[j], threadIdx.x, 128);
                    __syncthreads();
                    if (threadIdx.x == 0)
                        zp_stage0_done(a, (f << 8) | j, 1);
                }
            }
        });
    };
    /* the early pass: raw / RLE blocks at the positions stage 1 guessed, before anything is measured */
    if (early_pass)
        emu::launch(dim3(3), dim3(128), 64, [&]() {
            uint32_t *pos = reinterpret_cast<uint32_t *>(CRYO_SMEM_BASE());

            for (uint32_t f = a.n - a.early_frames + blockIdx.x; f < a.n; f += 3)
            {
                __syncthreads();
                if (threadIdx.x < 32)
                {
                    const uint32_t at = zp_frame_positions_warp<true>(a, f, threadIdx.x);

                    if (threadIdx.x < ZP_MAXB)
                        pos[threadIdx.x] = at;
                }
                __syncthreads();
                for (uint32_t j = 0; j < fr[(size_t) f * ZP_FF]; j++)
                    if (pos[j] != ~0u)
                        zp_stage0(a, (f << 8) | j, pos[j], threadIdx.x, 128);
            }
        });
    emu::launch(dim3(ngroups * split), dim3(32 * ZP2A_WARPS), ZP2A_SMEM, [&]() {
        ZP_FOR_GROUP_BLOCKS_CTA(a, blockIdx.x, split, ZPF_HUFMASK, threadIdx.x, zp_stage2a(a, g, j, CRYO_SMEM_BASE(), threadIdx.x));
    });
    emu::launch(dim3(ngroups * split), dim3(32), ZP2B_SMEM, [&]() {
        ZP_FOR_GROUP_BLOCKS(a, blockIdx.x, split, ZPF_HUFMASK, threadIdx.x, zp_stage2b(a, g, j, CRYO_SMEM_BASE(), threadIdx.x));
    });
    emu::launch(dim3(ngroups * split), dim3(32 * ZP3A_WARPS), ZP3A_SMEM, [&]() {
        ZP_FOR_GROUP_BLOCKS_CTA(a, blockIdx.x, split, ZPF_SEQMASK, threadIdx.x, zp_stage3a(a, g, j, CRYO_SMEM_BASE(), threadIdx.x));
    });
    emu::launch(dim3(ngroups * split), dim3(32),
                ZP3B_SMEM(ZP3B_SMALL, ZP3B_SMALL_LANES), [&]() {
        ZP_FOR_GROUP_BLOCKS(a, blockIdx.x, split, ZPF_SEQMASK, threadIdx.x,
                            (zp_stage3b<ZP3B_SMALL, 0, ZP3B_SMALL_LANES>(a, g, j, CRYO_SMEM_BASE(), threadIdx.x)));
    });
    emu::launch(dim3(ngroups * split), dim3(32), ZP3B_SMEM(ZP3B_LARGE, ZP_G), [&]() {
        ZP_FOR_GROUP_BLOCKS(a, blockIdx.x, split, ZPF_SEQMASK, threadIdx.x,
                            (zp_stage3b<ZP3B_LARGE, ZP3B_SMALL, ZP_G>(a, g, j, CRYO_SMEM_BASE(), threadIdx.x)));
    });
    if (!late_prefill)
        prefill();
    reinterpret_cast<uint32_t *>(ctl)[ZPC_SERVERS] = getenv("ZP_EMU_NO_SERVER") ? 0u : 1u;     /* stage 0's late pass is "running" */
    emu::launch(dim3(((unsigned) n + ZP4_WARPS - 1) / ZP4_WARPS), dim3(ZP4_THREADS), ZP4_SMEM, [&]() {
        const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        zp_stage4(a, blockIdx.x * ZP4_WARPS + warp, CRYO_SMEM_BASE() + warp * ZP4_PER_WARP, lane);
    });
    /* the runs stage 4 handed over: served here AFTER it (so a frame whose matches read such a run has waited in vain
     * and written it itself), then the check of what stage 0 published */
    if (a.jobs)
    {
        const uint32_t njobs = reinterpret_cast<uint32_t *>(ctl)[ZPC_JOB_TAIL];

        emu::launch(dim3(1), dim3(128), 0, [&]() {
            for (uint32_t k = 0; k < njobs; k++)
            {
                const uint32_t *job = jobs.data() + 4 * (size_t) k;

                if (!(job[3] & ZP_JOB_READY))
                    abort();
                team_fill_byte(a.dst + (size_t) job[0] * a.dst_stride + job[1], (uint8_t) job[3], job[2], threadIdx.x, 128);
                __syncthreads();
                if (threadIdx.x == 0 && !getenv("ZP_EMU_DROP_JOBS"))
                    jobs[4 * (size_t) k + 3] |= ZP_JOB_DONE;
            }
        });
    }
    if (late_prefill)
        prefill();
    if (a.jobs)
        emu::launch(dim3(((unsigned) n * ZP_JOBS + 255) / 256), dim3(256), 0, [&]() { zp_stage5_check(a, blockIdx.x * 256 + threadIdx.x); });
    if (getenv("ZP_DEBUG"))
        for (int i = 0; i < n; i++)
            for (uint32_t j = 0; j < fr[(size_t) i * ZP_FF]; j++)
            {
                const uint32_t *b = &blk[((size_t) i * ZP_MAXB + j) * ZP_BF];
                if ((b[ZPB_KIND] & 3) == 2)
                    fprintf(stderr, "frame %d block %u: lit_type %u regen %u huf_log %u nseq %u logs ll %u of %u ml %u\n", i, j,
                            (b[ZPB_KIND] >> 2) & 3, b[ZPB_REGEN], b[ZPB_HINFO] & 0xFF, b[ZPB_NSEQ], b[ZPB_SLOGS] & 0xFF,
                            (b[ZPB_SLOGS] >> 8) & 0xFF, (b[ZPB_SLOGS] >> 16) & 0x7F);
            }
    for (int i = 0; i < n; i++)
    {
        flags[i] = flag[i];
        if (!flag[i])
            continue;
        emu::launch(dim3(1), dim3(32), ZSW_PER_WARP, [&]() {
            zstdw_decode_frame(ib + off[i], csizes[i], o + i * (size_t) stride, cap, out_sizes + i, statuses + i,
                               sc, predef, CRYO_SMEM_BASE(), threadIdx.x);
        });
    }
    for (int i = 0; i < n; i++)
        memcpy(dsts[i], o + i * (size_t) stride, cap);
    for (int i = 1; i <= 64; i++)
        if (o[-i] != 0xAA)
            return -100;
    return 0;
}
