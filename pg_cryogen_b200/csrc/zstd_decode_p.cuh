/*
 * zstd_decode_p.cuh -- batched zstd frame decompression, PHASE-SPLIT (the default path).
 *
 * Replaces ZSTD_decompress as called at reference compression.c:116, like
 * zstd_decode_w.cuh (one warp per frame), whose building blocks it reuses.  What changes is
 * where the serial chains of the format run.  In the warp-per-frame decoder every serial
 * chain (FSE state walk, Huffman streams, weight decoding) is executed by a whole warp with
 * 1 or 4 useful lanes; ncu showed the kernel bound by instruction issue, not by HBM.  Here
 * each serial chain gets ONE LANE and 32 chains of different frames / zstd blocks run in
 * lockstep, and only the byte-moving part keeps a warp per frame:
 *
 *   stage 1  parse      lane per frame   frame + block headers -> block descriptors
 *   stage 2  literals   warp per (8 frames x block index): 8 Huffman tables in shared memory
 *                       (weights decoded one lane per table, table filled by the warp),
 *                       then 32 lanes = 8 blocks x 4 streams decode into the literal area
 *   stage 3  sequences  warp per (8 frames x block index): 24 FSE tables in shared memory
 *                       (counts read one lane per block, built by the warp), then one lane
 *                       per block walks the sequence bitstream -> (ll, ml, offset value)
 *   stage 4  execute    warp per frame: literal copies, matches, RLE / raw blocks through
 *                       the ring executor (cryo_wexec.cuh)
 *
 * The zstd blocks of one frame are entropy-decoded concurrently (their literals and
 * sequence streams do not depend on earlier output), which is 8x more parallelism per frame
 * than the format's sequential reading order suggests.
 *
 * The pipeline handles what libzstd writes for cryo blocks; everything else (skippable or
 * concatenated frames, more than ZP_MAXB blocks, Repeat_Mode tables, any malformed input,
 * pool exhaustion) raises the frame's flag and the warp-per-frame decoder decodes that frame
 * from scratch afterwards, so acceptance rules and status codes are its own in every case.
 */
#pragma once
#include "zstd_decode_w.cuh"

#define ZP_METHOD_ZSTD 1
#define ZP_MAXB   16u           /* zstd blocks per frame the pipeline takes (1 MiB / 128 KiB = 8) */
#define ZP_G      8u            /* frames per entropy warp */
#define ZP_BF     20u           /* u32 fields per block descriptor */
#define ZP_PREFILL_MIN 2048u     /* raw / RLE blocks at least this long are written ahead by stage 0 */
#define ZP_FF     6u            /* u32 fields per frame descriptor: nblk, fcs, has_fcs, route (1: stage 4 by a CTA, zstd_decode_c.cuh),
                                 * mask of the blocks with Huffman-coded literals, mask of the blocks with sequences */
/*
 * Long runs handed from stage 4 to stage 0 (round 2): the ~120 KB zero-run matches of a sparse frame, 96 % of the bytes
 * the warp executor wrote, stalled it for a quarter of its time beside the raw / RLE stage (profiles/r02_ablation.txt);
 * stage 0 has the bulk-copy engine.  A job is (frame, position, length, byte | ZP_JOB_READY), queued by stage 4 and
 * taken by stage 0's CTAs once their frames are done.  A job is published in its own slot (ZP_JOB_DONE), not in
 * pf_done: that counter says how many of a frame's BLOCKS stage 0 has written, in block order, and jobs finish in any
 * order relative to them (counted together, the later blocks of a frame passed for its jobs: a copy from a run not
 * yet written read whatever was there -- found by tests/test_emu_kernels.py's blocks with copies around long runs).
 */
#define ZP_JOBS        2u       /* per frame */
#define ZP_JOB_MIN     32768u   /* bytes */
#define ZP_JOB_READY   0x100u   /* word 3 of a job: the byte | READY once stage 4 has filled the slot in | DONE once stage 0 has written the run */
#define ZP_JOB_DONE    0x200u
#define ZPC_JOB_TAIL   8u       /* u32 indices into the control words at seq_alloc: jobs queued, */
#define ZPC_JOB_HEAD   9u       /* tickets taken, */
#define ZPC_EXEC_DONE  10u      /* warps of k_zp_execute that have finished, */
#define ZPC_EXEC_UP    11u      /* non-zero once k_zp_execute has started, */
#define ZPC_SERVERS    12u      /* CTAs of stage 0's late pass that are running: stage 4 queues a job only while there are some
                                 * (under a tool that runs the kernels one after the other there are none: it then writes its
                                 * runs itself, as in round 1, and stage 0 does not wait for jobs that cannot come) */
#define ZPB_RLEBYTE ZPB_LHDR     /* RLE blocks: the byte of the run (stage 4 then need not read the frame for it) */
#define ZPF_HUFMASK 4u
#define ZPF_SEQMASK 5u
#define ZP_CX_SEQS 8192u        /* frames with at least this many sequences take the CTA-per-frame stage 4 */

enum
{
    ZPB_OFF = 0,                /* frame-relative offset of the block content */
    ZPB_BSIZE,                  /* content bytes (raw, compressed) or run length (RLE) */
    ZPB_KIND,                   /* bits 0-1 block type, 2-3 literal type, 4 four streams, 8-15 modes */
    ZPB_LHDR,                   /* literals section header bytes (an RLE block: its byte, ZPB_RLEBYTE) */
    ZPB_REGEN,                  /* literals regenerated size */
    ZPB_LCSIZE,                 /* literals compressed size (tree description included) */
    ZPB_HDOFF,                  /* frame-relative offset of the Huffman tree description in force */
    ZPB_HDLEFT,                 /* bytes readable there */
    ZPB_NSEQ,
    ZPB_SEQOFF,                 /* block-relative offset of the first table description */
    ZPB_LITPOS,                 /* byte offset in the frame's literal area (multiple of 16) */
    ZPB_SEQPOS,                 /* entry offset in the frame's sequence area */
    ZPB_SPECPOS,                /* raw / RLE block that stage 0 is to write: 0, any other block: ~0u */
    ZPB_HDBLK,                  /* index of the block whose Huffman tree is in force */
    ZPB_HINFO,                  /* written by stage 2a: table log | description bytes << 8 (0: failed) */
    ZPB_SLOGS,                  /* written by stage 3a: ll_log | of_log << 8 | ml_log << 16 | 1 << 31 */
    ZPB_BITOFF,                 /* written by stage 3a: block-relative offset of the sequence bitstream */
    ZPB_OUTSZ,                  /* bytes the block regenerates: stage 1 (raw, RLE, no sequences) or stage 3b (literals + matches) */
    ZPB_SPECAT                  /* raw / RLE block that stage 0 may write EARLY: its output position if every Compressed block
                                 * before it regenerates a full 128 KiB (what libzstd cuts), ~0u: no such guess */
};

struct ZpArgs
{
    const int32_t  *methods;
    const uint8_t  *src;
    const uint64_t *src_off;
    const uint32_t *src_size;
    uint8_t        *dst;
    uint64_t        dst_stride;
    uint32_t        cap;
    uint32_t        n;
    uint32_t       *out_size;
    int32_t        *status;
    uint32_t       *fr;         /* n x ZP_FF */
    uint32_t       *blk;        /* n x ZP_MAXB x ZP_BF */
    uint32_t       *flag;       /* n; non-zero: frame goes to the warp-per-frame decoder */
    uint64_t       *seqbase;    /* n; first entry of the frame in seq[] */
    unsigned long long *seq_alloc;
    uint32_t       *cxlist;     /* frames routed to the CTA-per-frame stage 4, in no particular order */
    uint32_t       *cxcount;    /* how many */
    uint32_t       *pf_done;    /* n; blocks of the frame stage 0 has finished (release / acquire with stage 4) */
    uint32_t        early_frames; /* the early pass of stage 0 takes the last early_frames frames of the batch (0: there is none) */
    uint32_t        pf_hint;    /* bit 0: stage 0 bulk stores with the L2 evict_first policy; bit 1: stage 4 asks L2 for
                                 * sequences and literals a few loads ahead; bit 2: every frame takes the CTA-per-frame
                                 * stage 4 (small batches); bit 3: none does; bit 4: stage 1 asks the frames into L2 */
    uint32_t       *jobs;       /* n x ZP_JOBS x 4 words, zeroed per call; nullptr: stage 4 writes its long runs itself */
    uint32_t       *pf_expect;  /* n; blocks + jobs stage 4 left to stage 0 (checked against pf_done after both) */
    uint32_t        exec_warps; /* warps of the k_zp_execute launch */
    uint32_t        pf_inflight; /* stage 0: 128 KiB bulk groups a CTA keeps in flight (head-of-line blocking, see k_zp_prefill) */
    uint8_t        *lit;        /* n x lit_stride: Huffman-decoded literals */
    uint64_t        lit_stride; /* multiple of 16, >= cap + 16 * ZP_MAXB */
    uint64_t       *seq;        /* ll | ml << 17 | offset_value << 35 */
    uint64_t        seq_cap;
    const uint32_t *predef;
    uint16_t       *huftab;     /* n x ZP_MAXB x u16[2048] */
    uint32_t       *fsetab;     /* n x ZP_MAXB x u32[ZP3_CELLS] */
};

/* bytes of device memory behind ZpArgs for n frames of capacity cap */
static inline uint64_t zp_lit_stride(uint32_t cap) { return ((uint64_t) cap + 16u * ZP_MAXB + 15u) & ~15ull; }
static inline uint64_t zp_seq_cap(uint64_t n, uint32_t cap) { return n * (uint64_t) (cap / 8u + 64u); }

/* ------------------------------------------------------------------ stage 1: parse ---- */

/* One lane, one frame.  False: not for the pipeline (the flag is raised by the caller). */
CRYO_DEV bool zp_parse(const uint8_t *in, uint32_t csize, uint32_t cap, uint64_t lit_stride,
                       uint32_t *fr, uint32_t *blk, uint32_t *seq_total_out)
{
    if (csize < 6)
        return false;
    const uint32_t magic = in[0] | ((uint32_t) in[1] << 8) | ((uint32_t) in[2] << 16) | ((uint32_t) in[3] << 24);

    if (magic != 0xFD2FB528u)
        return false;
    uint32_t ip = 4;
    const uint32_t fhd = in[ip++];
    const uint32_t fcs_flag = fhd >> 6, single = (fhd >> 5) & 1u, checksum = (fhd >> 2) & 1u;
    const uint32_t dict_flag = fhd & 3u;

    if (fhd & 0x08u)
        return false;
    if (!single)
    {
        const uint32_t b = in[ip++];

        if (10 + (b >> 3) > 27)
            return false;
    }
    const uint32_t dict_bytes = dict_flag == 3 ? 4 : dict_flag;
    const uint32_t fcs_bytes = fcs_flag == 0 ? (single ? 1u : 0u) : (1u << fcs_flag);

    if (ip + dict_bytes + fcs_bytes > csize)
        return false;
    for (uint32_t i = 0; i < dict_bytes; i++)
        if (in[ip + i] != 0)
            return false;                       /* no dictionary on this path */
    ip += dict_bytes;
    uint64_t fcs = 0;

    for (uint32_t i = 0; i < fcs_bytes; i++)
        fcs |= (uint64_t) in[ip + i] << (8 * i);
    if (fcs_bytes == 2)
        fcs += 256;
    ip += fcs_bytes;
    if (fcs_bytes && fcs > cap)
        return false;

    uint32_t nb = 0, lit_total = 0, seq_total = 0, hd_off = 0, hd_left = 0, hd_blk = 0;
    uint32_t hufmask = 0, seqmask = 0;
    uint64_t spec_at = 0;                       /* where the block would start if the Compressed ones before it were full */
    bool     have_hd = false;

    for (;;)
    {
        if (ip + 3 > csize || nb >= ZP_MAXB)
            return false;
        const uint32_t bh = in[ip] | ((uint32_t) in[ip + 1] << 8) | ((uint32_t) in[ip + 2] << 16);
        const uint32_t last = bh & 1u, type = (bh >> 1) & 3u, bsize = bh >> 3;
        uint32_t *b = blk + nb * ZP_BF;

        ip += 3;
        if (type == 3 || bsize > ZS_MAXBLOCK)
            return false;
        b[ZPB_OFF] = ip;
        b[ZPB_BSIZE] = bsize;
        b[ZPB_KIND] = type;
        b[ZPB_SPECPOS] = (type < 2 && bsize >= ZP_PREFILL_MIN) ? 0u : ~0u;
        b[ZPB_OUTSZ] = bsize;                   /* raw, RLE; Compressed: below and stage 3b */
        /* a guess is only made inside a declared content size: a frame that decodes rewrites all of it */
        b[ZPB_SPECAT] = (type < 2 && bsize >= ZP_PREFILL_MIN && fcs_bytes && spec_at + bsize <= fcs) ? (uint32_t) spec_at : ~0u;
        spec_at += type < 2 ? bsize : ZS_MAXBLOCK;
        if (type == 0)
        {
            if (bsize > csize - ip)
                return false;
            ip += bsize;
        }
        else if (type == 1)
        {
            if (ip + 1 > csize)
                return false;
            b[ZPB_RLEBYTE] = in[ip];
            ip += 1;
        }
        else
        {
            if (bsize == 0 || bsize > csize - ip)
                return false;
            const uint8_t *bp = in + ip;
            const uint32_t lt = bp[0] & 3u, sf = (bp[0] >> 2) & 3u;
            uint32_t lhdr, regen, lcsize, four = 0;

            if (lt < 2)
            {
                if (sf == 0 || sf == 2)
                {
                    lhdr = 1;
                    regen = bp[0] >> 3;
                }
                else if (sf == 1)
                {
                    if (bsize < 2)
                        return false;
                    lhdr = 2;
                    regen = (bp[0] >> 4) | ((uint32_t) bp[1] << 4);
                }
                else
                {
                    if (bsize < 3)
                        return false;
                    lhdr = 3;
                    regen = (bp[0] >> 4) | ((uint32_t) bp[1] << 4) | ((uint32_t) bp[2] << 12);
                }
                lcsize = lt == 0 ? regen : 1;
            }
            else
            {
                if (bsize < 5)
                    return false;
                const uint64_t v = bp[0] | ((uint64_t) bp[1] << 8) | ((uint64_t) bp[2] << 16) |
                                   ((uint64_t) bp[3] << 24) | ((uint64_t) bp[4] << 32);

                if (sf < 2)
                {
                    lhdr = 3;
                    regen = (uint32_t) (v >> 4) & 0x3FFu;
                    lcsize = (uint32_t) (v >> 14) & 0x3FFu;
                    four = sf == 0 ? 0 : 1;
                }
                else if (sf == 2)
                {
                    lhdr = 4;
                    regen = (uint32_t) (v >> 4) & 0x3FFFu;
                    lcsize = (uint32_t) (v >> 18) & 0x3FFFu;
                    four = 1;
                }
                else
                {
                    lhdr = 5;
                    regen = (uint32_t) (v >> 4) & 0x3FFFFu;
                    lcsize = (uint32_t) (v >> 22) & 0x3FFFFu;
                    four = 1;
                }
            }
            if (lhdr + lcsize > bsize || regen > ZS_MAXBLOCK)
                return false;
            if (lt == 2)
            {
                have_hd = true;
                hd_off = ip + lhdr;
                hd_left = lcsize;
                hd_blk = nb;
            }
            else if (lt == 3 && !have_hd)
                return false;
            /* sequences section header */
            uint32_t sp = lhdr + lcsize, nseq, modes = 0;

            if (sp + 1 > bsize)
                return false;
            if (bp[sp] < 128)
            {
                nseq = bp[sp];
                sp += 1;
            }
            else if (bp[sp] < 255)
            {
                if (sp + 2 > bsize)
                    return false;
                nseq = ((uint32_t) (bp[sp] - 128) << 8) + bp[sp + 1];
                sp += 2;
            }
            else
            {
                if (sp + 3 > bsize)
                    return false;
                nseq = bp[sp + 1] + ((uint32_t) bp[sp + 2] << 8) + 0x7F00u;
                sp += 3;
            }
            if (nseq)
            {
                if (sp + 1 > bsize)
                    return false;
                modes = bp[sp++];
                /* reserved bits; Repeat_Mode tables are left to the warp-per-frame decoder */
                if ((modes & 3u) || (modes >> 6) == 3u || ((modes >> 4) & 3u) == 3u || ((modes >> 2) & 3u) == 3u)
                    return false;
            }
            else if (sp != bsize)
                return false;
            b[ZPB_KIND] = type | (lt << 2) | (four << 4) | (modes << 8);
            b[ZPB_LHDR] = lhdr;
            b[ZPB_REGEN] = regen;
            b[ZPB_LCSIZE] = lcsize;
            b[ZPB_HDOFF] = hd_off;
            b[ZPB_HDLEFT] = hd_left;
            b[ZPB_HDBLK] = hd_blk;
            b[ZPB_HINFO] = 0;
            b[ZPB_SLOGS] = 0;
            b[ZPB_BITOFF] = 0;
            b[ZPB_OUTSZ] = regen;                /* + the match lengths, added by stage 3b */
            b[ZPB_NSEQ] = nseq;
            b[ZPB_SEQOFF] = sp;
            b[ZPB_LITPOS] = lit_total;
            b[ZPB_SEQPOS] = seq_total;
            if (lt >= 2)
            {
                lit_total += (regen + 15u) & ~15u;
                hufmask |= 1u << nb;
            }
            if (nseq)
                seqmask |= 1u << nb;
            seq_total += nseq;
            if ((uint64_t) lit_total > lit_stride)
                return false;
            ip += bsize;
        }
        nb++;
        if (last)
            break;
    }
    if (checksum)
    {
        if (ip + 4 > csize)
            return false;
        ip += 4;                                /* XXH64 content checksum: skipped, not verified */
    }
    if (ip != csize)
        return false;                           /* another frame follows: not for the pipeline */
    fr[0] = nb;
    fr[1] = (uint32_t) fcs;
    fr[2] = fcs_bytes ? 1u : 0u;
    fr[ZPF_HUFMASK] = hufmask;
    fr[ZPF_SEQMASK] = seqmask;
    *seq_total_out = seq_total;
    return true;
}

/* stage 1 body: the calling lane owns frame f */
CRYO_DEV void zp_stage1(const ZpArgs &a, uint32_t f)
{
    uint32_t *fr = a.fr + (size_t) f * ZP_FF;

    fr[0] = 0;
    fr[ZPF_HUFMASK] = 0;
    fr[ZPF_SEQMASK] = 0;
    a.flag[f] = 0;
    a.pf_done[f] = 0;
    if (a.pf_expect)
        a.pf_expect[f] = 0;
    if (a.methods[f] != ZP_METHOD_ZSTD)
        return;
    uint32_t seq_total = 0;
    bool     ok = zp_parse(a.src + a.src_off[f], a.src_size[f], a.cap, a.lit_stride, fr,
                           a.blk + (size_t) f * ZP_MAXB * ZP_BF, &seq_total);

    if (ok)
    {
        const unsigned long long base = atomicAdd(a.seq_alloc, (unsigned long long) seq_total);

        if (base + seq_total > a.seq_cap)
            ok = false;
        a.seqbase[f] = base;
        /* who executes the frame: a warp (k_zp_execute, with stage 0 writing its Raw / RLE blocks), or a
         * whole CTA (k_zp_execute_c), which writes every block itself */
        const bool cx = !(a.pf_hint & 8u) && ((a.pf_hint & 4u) || seq_total >= ZP_CX_SEQS);

        fr[3] = cx ? 1u : 0u;
        if (cx)
        {
            for (uint32_t j = 0; j < fr[0]; j++)
            {
                a.blk[((size_t) f * ZP_MAXB + j) * ZP_BF + ZPB_SPECPOS] = ~0u;
                a.blk[((size_t) f * ZP_MAXB + j) * ZP_BF + ZPB_SPECAT] = ~0u;
            }
            a.cxlist[atomicAdd(a.cxcount, 1u)] = f;
        }
    }
    if (!ok)
    {
        fr[0] = 0;
        fr[ZPF_HUFMASK] = 0;
        fr[ZPF_SEQMASK] = 0;
        a.flag[f] = 1;
    }
}

/*
 * How the lane-serial stages (2b, 3b) are laid over the grid.  A warp of those stages takes one block
 * index of ZP_G consecutive frames.  Round 1 launched one warp per (group, block index), sixteen per
 * group: on the sparse headline table fourteen of the sixteen found nothing and left, and the 862
 * literal warps that had work did not fit the 740 the shared memory admits at once -- the stage ran
 * as two waves, the second one a sixth full, and took twice the time of its longest chain.  Now a
 * warp walks the block indices that have work (the union of its frames' masks, from stage 1) itself,
 * and `split` warps share the indices of a group by rank; the host picks split so that the grid fits
 * the device in one wave when it can.
 */
CRYO_DEV uint32_t zp_group_mask(const ZpArgs &a, uint32_t g, uint32_t field, uint32_t lane)
{
    const uint32_t f = g * ZP_G + (lane & (ZP_G - 1u));
    const uint32_t m = (f < a.n && a.fr[(size_t) f * ZP_FF] != 0) ? a.fr[(size_t) f * ZP_FF + field] : 0u;

    return __reduce_or_sync(CRYO_FULL, m);
}
/* the table stages (2a, 3a) take a group with a whole CTA: the same walk with a barrier between the calls */
#define ZP_FOR_GROUP_BLOCKS_CTA(a, w, split, field, tid, call)                               \
    {                                                                                        \
        const uint32_t g_ = (w) / (split), c_ = (w) % (split);                               \
        uint32_t m_ = zp_group_mask((a), g_, (field), (tid) & 31u);                          \
                                                                                             \
        for (uint32_t k_ = 0; m_; m_ &= m_ - 1u, k_++)                                       \
        {                                                                                    \
            if (k_ % (split) != c_)                                                          \
                continue;                                                                    \
            const uint32_t g = g_, j = (uint32_t) __ffs((int) m_) - 1u;                      \
                                                                                             \
            call;                                                                            \
            __syncthreads();                                                                 \
        }                                                                                    \
    }
#define ZP_FOR_GROUP_BLOCKS(a, w, split, field, lane, call)                                  \
    {                                                                                        \
        const uint32_t g_ = (w) / (split), c_ = (w) % (split);                               \
        uint32_t m_ = zp_group_mask((a), g_, (field), (lane));                               \
                                                                                             \
        for (uint32_t k_ = 0; m_; m_ &= m_ - 1u, k_++)                                       \
        {                                                                                    \
            if (k_ % (split) != c_)                                                          \
                continue;                                                                    \
            const uint32_t g = g_, j = (uint32_t) __ffs((int) m_) - 1u;                      \
                                                                                             \
            call;                                                                            \
            __syncwarp();                                                                    \
        }                                                                                    \
    }

/* ------------------------------------------------- stage 0: raw / RLE blocks ahead ---- */

/*
 * The output position of a Raw or RLE block is known once the blocks before it have been
 * measured (stage 1 for raw, RLE and sequence-less blocks; stage 3b sums the match lengths of
 * the others; zp_frame_positions).  Stage 0 writes those blocks, at HBM speed, WHILE stage 4
 * runs (that one is bound by instruction issue); stage 4 steps over them.  It needs their bytes
 * only when a later match reads them: before the first such read it waits until
 * pf_done[frame] says stage 0 has published the frame's blocks; should that take too long
 * (stage 0 not scheduled yet) it writes them itself, which is idempotent, so no ordering
 * between the two kernels is assumed.  One persistent CTA per SM takes the frames in index
 * order (item = frame << 8 | block).
 */
/*
 * Output positions of frame f's blocks for stage 0, from the regenerated sizes (stage 1 for raw,
 * RLE and sequence-less blocks, stage 3b for the others): pos[j] = position of block j if stage 0
 * is to write it, ~0u otherwise (not a candidate, would not fit the capacity, frame flagged).
 * Stage 4 arrives at the same positions by construction: it regenerates exactly those sizes.
 */
CRYO_DEV void zp_frame_positions(const ZpArgs &a, uint32_t f, uint32_t *pos)
{
    const uint32_t nb = a.fr[(size_t) f * ZP_FF];
    const bool     live = a.flag[f] == 0;
    uint64_t at = 0;

    for (uint32_t j = 0; j < ZP_MAXB; j++)
    {
        const uint32_t *b = a.blk + ((size_t) f * ZP_MAXB + j) * ZP_BF;

        pos[j] = ~0u;
        if (j >= nb)
            continue;
        if (live && b[ZPB_SPECPOS] != ~0u && at + b[ZPB_BSIZE] <= a.cap)
            pos[j] = (uint32_t) at;
        at += b[ZPB_OUTSZ];
    }
}

/* did the early pass of stage 0 (raw / RLE blocks at their guessed positions, ZPB_SPECAT) take frame f */
CRYO_DEV bool zp_early_frame(const ZpArgs &a, uint32_t f) { return f + a.early_frames >= a.n; }

/* the late pass of stage 0: drop the blocks the early pass wrote at the right place (stage 4 applies the same
 * test when it steps over them, so the two agree on how many blocks pf_done counts) */
CRYO_DEV void zp_frame_positions_late(const ZpArgs &a, uint32_t f, uint32_t *pos)
{
    zp_frame_positions(a, f, pos);
    if (zp_early_frame(a, f))
        for (uint32_t j = 0; j < ZP_MAXB; j++)
            if (pos[j] != ~0u && a.blk[((size_t) f * ZP_MAXB + j) * ZP_BF + ZPB_SPECAT] == pos[j])
                pos[j] = ~0u;
}

/*
 * The same by one warp, lane = block (what the stage itself uses: a thread walking the sixteen descriptors
 * one after the other took 10 us per frame, most of the 26 us a CTA spent on a sparse frame).  EARLY: the
 * guessed positions; otherwise the exact ones without the blocks the early pass placed right.
 */
template <bool EARLY>
CRYO_DEV uint32_t zp_frame_positions_warp(const ZpArgs &a, uint32_t f, uint32_t lane)
{
    const uint32_t nb = a.fr[(size_t) f * ZP_FF];
    const uint32_t *b = a.blk + ((size_t) f * ZP_MAXB + (lane & (ZP_MAXB - 1u))) * ZP_BF;
    const bool     have = lane < ZP_MAXB && lane < nb;

    if (EARLY)
        return have ? b[ZPB_SPECAT] : ~0u;
    const bool     live = a.flag[f] == 0;
    const uint32_t bsize = have ? b[ZPB_BSIZE] : 0u, specpos = have ? b[ZPB_SPECPOS] : ~0u;
    const uint32_t specat = have ? b[ZPB_SPECAT] : ~0u;
    unsigned long long at = have ? b[ZPB_OUTSZ] : 0u;

#pragma unroll
    for (uint32_t d = 1; d < ZP_MAXB; d <<= 1)
    {
        const unsigned long long up = __shfl_up_sync(CRYO_FULL, at, d);

        if (lane >= d)
            at += up;
    }
    at -= have ? b[ZPB_OUTSZ] : 0u;             /* exclusive */
    uint32_t pos = (have && live && specpos != ~0u && at + bsize <= a.cap) ? (uint32_t) at : ~0u;

    if (zp_early_frame(a, f) && pos == specat)
        pos = ~0u;
    return pos;
}

CRYO_DEV void zp_stage0(const ZpArgs &a, uint32_t item, uint32_t at, uint32_t tid, uint32_t nthr)
{
    const uint32_t f = item >> 8, j = item & 0xFFu;
    const uint32_t *b = a.blk + ((size_t) f * ZP_MAXB + j) * ZP_BF;
    const uint8_t *in = a.src + a.src_off[f] + b[ZPB_OFF];
    uint8_t *dst = a.dst + (size_t) f * a.dst_stride + at;

    if ((b[ZPB_KIND] & 3u) == 0)
        team_copy(dst, in, b[ZPB_BSIZE], tid, nthr);
    else
        team_fill_byte(dst, in[0], b[ZPB_BSIZE], tid, nthr);
}

/* after every thread of the CTA has finished (and fenced) `blocks` blocks of item's frame: one thread publishes them */
CRYO_DEV void zp_stage0_done(const ZpArgs &a, uint32_t item, uint32_t blocks)
{
    __threadfence();
    atomicAdd(a.pf_done + (item >> 8), blocks);
}

CRYO_DEV void zp_st_release(uint32_t *p, uint32_t v)
{
#ifdef CRYO_EMU
    *reinterpret_cast<volatile uint32_t *>(p) = v;
#else
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
}

CRYO_DEV uint32_t zp_ld_acquire(const uint32_t *p)
{
#ifdef CRYO_EMU
    return *reinterpret_cast<const volatile uint32_t *>(p);
#else
    uint32_t v;

    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
#endif
}

CRYO_DEV void zp_prefetch_l2(const uint8_t *p)
{
#ifndef CRYO_EMU
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void) p;
#endif
}

/* --------------------------------------------------------------- stage 2: literals ---- */

/*
 * 2a: one CTA per (ZP_G frames x block index): the weights of the trees decoded one lane per
 *     tree, then each table filled by a warp straight into the block's slot of huftab (global,
 *     u16[2048] per slot); log and description length go into the block descriptor.
 * 2b: one warp per (8 frames x block index), 32 lanes = 8 blocks x 4 streams, tables read
 *     through L1.  No shared memory: the stage co-resides with anything.
 */
#define ZP2A_WARPS      4u
#define ZP2A_LW         708u                    /* per tree: weights 256 | wfse 256 | wcounts 32 | wnext 128; odd word stride */
#define ZP2A_DESC       176u                    /* staged tree description: 129 bytes at most + alignment slack */
#define ZP2A_OFF_DESC   (ZP_G * ZP2A_LW)        /* 5 664: multiple of 16 */
#define ZP2A_OFF_WORK   (ZP2A_OFF_DESC + ZP_G * ZP2A_DESC)      /* per warp: symstart u16[256] | rankc u32[32] */
#define ZP2A_OFF_META   (ZP2A_OFF_WORK + ZP2A_WARPS * 640u)
#define ZP2A_SMEM       (ZP2A_OFF_META + 128u)

/* Huffman weights of one tree description, one lane (RFC 8878 4.2.1).  Bytes used or 0. */
CRYO_DEV uint32_t zp_huf_weights(const uint8_t *src, uint32_t n, uint8_t *weights, uint32_t *wfse,
                                 int16_t *wcounts, uint16_t *wnext, uint32_t *nw_out)
{
    if (n == 0)
        return 0;
    const uint32_t h = src[0];
    uint32_t nw = 0, used;

    if (h >= 128)
    {
        nw = h - 127;
        used = 1 + (nw + 1) / 2;
        if (used > n)
            return 0;
        for (uint32_t i = 0; i < nw; i++)
        {
            const uint32_t b = src[1 + i / 2];

            weights[i] = (uint8_t) ((i & 1) ? (b & 15u) : (b >> 4));
        }
    }
    else
    {
        used = 1 + h;
        if (used > n || h == 0)
            return 0;
        int32_t  nsym = 0, flog = 0;
        const uint32_t hdr = fse_read_counts(src + 1, h, 6, 12, wcounts, &nsym, &flog);
        BitsBack bb;

        if (hdr == 0 || hdr >= h)
            return 0;
        fse_build_table(wfse, wcounts, nsym, flog, wnext);
        if (!bb_init(bb, src + 1 + hdr, h - hdr))
            return 0;
        bb_refill(bb);
        uint32_t s1 = bb_read(bb, (uint32_t) flog);
        uint32_t s2 = bb_read(bb, (uint32_t) flog);

        for (;;)
        {
            if (nw > 253)
                return 0;
            const uint32_t c1 = wfse[s1];

            weights[nw++] = (uint8_t) c1;
            bb_refill(bb);
            s1 = (c1 >> 16) + bb_read(bb, (c1 >> 8) & 0xFFu);
            if (bb.remaining < 0)
            {
                weights[nw++] = (uint8_t) wfse[s2];
                break;
            }
            const uint32_t c2 = wfse[s2];

            weights[nw++] = (uint8_t) c2;
            bb_refill(bb);
            s2 = (c2 >> 16) + bb_read(bb, (c2 >> 8) & 0xFFu);
            if (bb.remaining < 0)
            {
                weights[nw++] = (uint8_t) wfse[s1];
                break;
            }
        }
    }
    *nw_out = nw;
    return used;
}

/*
 * stage 2a body: one CTA of ZP2A_WARPS warps, block index j of frames [g * ZP_G, g * ZP_G + ZP_G).
 * Phase A, warp 0: lane i decodes the weights of frame g * ZP_G + i's tree (a serial chain per
 * tree: ZP_G of them in lockstep instead of one per warp).  Phase B: one table per warp (with 32
 * frames per CTA, as in round 1, a warp filled four tables one after the other and phase B was
 * 57 % of the stage; profiles/r02n), filled straight into its global slot.  Called once per
 * block index the CTA takes (ZP_FOR_GROUP_BLOCKS_CTA): a barrier separates the calls.
 */
CRYO_DEV void zp_stage2a(const ZpArgs &a, uint32_t g, uint32_t j, uint8_t *smem, uint32_t tid)
{
    const uint32_t warp = tid >> 5, lane = tid & 31u;
    uint32_t *s_nw = reinterpret_cast<uint32_t *>(smem + ZP2A_OFF_META);        /* nw | used << 16, 0: nothing to build */

    if (warp == 0 && lane < ZP_G)
    {
        const uint32_t f = g * ZP_G + lane;
        uint32_t meta = 0;

        if (f < a.n && a.fr[(size_t) f * ZP_FF] > j)
        {
            const uint32_t *b = a.blk + ((size_t) f * ZP_MAXB + j) * ZP_BF;
            const uint32_t kind = b[ZPB_KIND];

            /* treeless blocks use the slot of the block that defined the tree */
            if ((kind & 3u) == 2u && ((kind >> 2) & 3u) == 2u)
            {
                /* the description (129 bytes at most) -> shared memory, then the serial part from there */
                const uint8_t *gd = a.src + a.src_off[f] + b[ZPB_HDOFF];
                const uint32_t dleft = b[ZPB_HDLEFT] < 130u ? b[ZPB_HDLEFT] : 130u;
                const uint32_t dd = (uint32_t) ((uintptr_t) gd & 15u);
                uint8_t *sd = smem + ZP2A_OFF_DESC + lane * ZP2A_DESC;
                uint8_t *lw = smem + lane * ZP2A_LW;
                uint32_t used, nw = 0;

                for (uint32_t k = 0; k < dd + dleft; k += 16)
                    st16(sd + k, ld16(gd - dd + k));
                used = zp_huf_weights(sd + dd, dleft, lw, reinterpret_cast<uint32_t *>(lw + 256),
                                      reinterpret_cast<int16_t *>(lw + 512), reinterpret_cast<uint16_t *>(lw + 544), &nw);
                if (used == 0)
                    a.flag[f] = 1;
                else
                    meta = nw | (used << 16);
            }
        }
        s_nw[lane] = meta;
    }
    __syncthreads();
    for (uint32_t t = warp; t < ZP_G; t += ZP2A_WARPS)
    {
        const uint32_t meta = s_nw[t];

        if (meta == 0)
            continue;
        const uint32_t f = g * ZP_G + t;
        int32_t    log = 0;
        const bool ok = zsw_huf_table(smem + t * ZP2A_LW, meta & 0xFFFFu, a.huftab + ((size_t) f * ZP_MAXB + j) * 2048u,
                                      reinterpret_cast<uint16_t *>(smem + ZP2A_OFF_WORK + warp * 640u),
                                      reinterpret_cast<uint32_t *>(smem + ZP2A_OFF_WORK + warp * 640u + 512u), &log, lane);

        if (lane == 0)
        {
            if (ok)
                a.blk[((size_t) f * ZP_MAXB + j) * ZP_BF + ZPB_HINFO] = (uint32_t) log | ((meta >> 16) << 8);
            else
                a.flag[f] = 1;
        }
        __syncwarp();
    }
}

/*
 * stage 2b body: one warp, block index j of frames [g * ZP_G, g * ZP_G + ZP_G); lane = 4 *
 * frame + stream.
 *
 * The 32 streams are decoded in lockstep, and the loop is written for that: the scoreboard
 * tracks registers per warp, so a load issued by one lane under a divergent branch delays the
 * next use of that register by every other lane.  Hence no branch and no global load in the
 * symbol loop: each lane's stream is staged through a 256-byte window in shared memory
 * (refilled by all lanes together when any lane runs low), the refill of the bit accumulator
 * is predicated, and the window word for the next refill is fetched right after the current
 * one so that its latency is covered by two symbol decodes.
 */
#define ZP2B_WIN        256u                    /* bytes of stream per lane window */
#define ZP2B_WSTRIDE    (ZP2B_WIN / 4u + 1u)    /* words; odd stride: conflict-free when lanes read the same offset */
#define ZP2B_SMEM       (ZP_G * 4096u + 32u * ZP2B_WSTRIDE * 4u)

/* (two symbols per table lookup was tried here: u32 entries rebuilt from the one-symbol table,
 * byte stores, data-dependent trip count.  On B200 it ran 2.5x slower than this loop on dense
 * hex text, profiles/r01e_arrangements.txt, and was dropped.) */
CRYO_DEV void zp_stage2b(const ZpArgs &a, uint32_t g, uint32_t j, uint8_t *smem, uint32_t lane)
{
    constexpr uint32_t TB = 4096u, OFF_WIN = ZP_G * TB;

    const uint32_t ti = lane >> 2, tf = g * ZP_G + ti, s = lane & 3u;
    const uint32_t *b = nullptr;
    uint32_t kind = 0, lt = 0, hb = 0, hinfo = 0;
    bool     valid = false;

    if (tf < a.n && a.fr[(size_t) tf * ZP_FF] > j && a.flag[tf] == 0)
    {
        b = a.blk + ((size_t) tf * ZP_MAXB + j) * ZP_BF;
        kind = b[ZPB_KIND];
        lt = (kind >> 2) & 3u;
        valid = (kind & 3u) == 2u && lt >= 2u;
        if (valid)
        {
            hb = b[ZPB_HDBLK];                  /* the block whose tree is in force (itself for lt == 2) */
            hinfo = a.blk[((size_t) tf * ZP_MAXB + hb) * ZP_BF + ZPB_HINFO];
        }
    }
    if (__ballot_sync(CRYO_FULL, valid) == 0)
        return;
    /* the group's tables -> shared memory, by the 4 lanes of each table */
    if (valid && (hinfo & 0xFFu) != 0)
    {
        const uint16_t *g1 = a.huftab + ((size_t) tf * ZP_MAXB + hb) * 2048u;
        const uint32_t log1 = hinfo & 0xFFu, size1 = 1u << log1;

        {
            const uint8_t *gt = reinterpret_cast<const uint8_t *>(g1);
            uint8_t *st = smem + ti * TB;
            const uint32_t bytes = 2u * size1;

            if (bytes < 16u)
            {
                if (s == 0)
                    for (uint32_t k = 0; k < bytes; k += 2)
                        *reinterpret_cast<uint16_t *>(st + k) = *reinterpret_cast<const uint16_t *>(gt + k);
            }
            else
                for (uint32_t k = 16u * s; k < bytes; k += 64u)
                    st16(st + k, ld16(gt + k));
        }
    }
    /* this lane's stream: src[0, sn) -> cnt symbols at dst */
    const uint8_t *src = nullptr;
    uint8_t  *dst = nullptr;
    uint32_t  sn = 0, cnt = 0;
    int32_t   tlog = 1;
    bool      ok = true, act = false;

    if (valid)
    {
        tlog = (int32_t) (hinfo & 0xFFu);
        const uint32_t own = lt == 2u ? hinfo >> 8 : 0u;
        const uint8_t *p = a.src + a.src_off[tf] + b[ZPB_OFF] + b[ZPB_LHDR] + own;
        const uint32_t left = b[ZPB_LCSIZE] - own, regen = b[ZPB_REGEN];
        uint8_t *d0 = a.lit + (size_t) tf * a.lit_stride + b[ZPB_LITPOS];

        if (tlog == 0)
        {
            ok = false;                         /* the tree's own stage-2a warp flagged the frame */
            tlog = 1;
        }
        else if (!((kind >> 4) & 1u))
        {
            if (s == 0)
            {
                src = p;
                sn = left;
                dst = d0;
                cnt = regen;
                act = true;
            }
        }
        else if (left < 6)
            ok = false;
        else
        {
            const uint32_t s1 = p[0] | ((uint32_t) p[1] << 8);
            const uint32_t s2 = p[2] | ((uint32_t) p[3] << 8);
            const uint32_t s3 = p[4] | ((uint32_t) p[5] << 8);
            const uint32_t seg = (regen + 3) / 4;

            if (6 + s1 + s2 + s3 > left || seg * 3 > regen)
                ok = false;
            else
            {
                const uint32_t s4 = left - 6 - s1 - s2 - s3;
                const uint32_t so = s == 0 ? 0 : s == 1 ? s1 : s == 2 ? s1 + s2 : s1 + s2 + s3;

                src = p + 6 + so;
                sn = s == 0 ? s1 : s == 1 ? s2 : s == 2 ? s3 : s4;
                dst = d0 + s * seg;
                cnt = s < 3 ? seg : regen - 3 * seg;
                act = true;
            }
        }
        if (act && sn == 0)
        {
            ok = false;
            act = false;
        }
    }
    /*
     * positions are byte offsets from abase = src rounded down to 16 (so they can go below zero:
     * words in front of the stream read as zero, a valid stream never consumes them and a
     * corrupt one fails the final bit count)
     */
    uint32_t *win = reinterpret_cast<uint32_t *>(smem + OFF_WIN) + lane * ZP2B_WSTRIDE;
    const uint8_t *abase = src - ((uintptr_t) src & 15u);
    const int32_t  delta = (int32_t) ((uintptr_t) src & 15u);
    int32_t  npos = act ? (int32_t) ((delta + sn - 1u) & ~3u) : 0;      /* offset of the next word to hand out */
    int32_t  g0 = 0;                            /* offset of the window's first byte, multiple of 16 */

#define ZP2B_FILL()                                                          \
    {                                                                        \
        __syncwarp();                                                        \
        g0 = (npos & ~15) - (int32_t) (ZP2B_WIN - 16u);                      \
        _Pragma("unroll") for (int r_ = 0; r_ < 2; r_++)                     \
        {                                                                    \
            uint4 v_[8];                                                     \
            _Pragma("unroll") for (int k_ = 0; k_ < 8; k_++)                 \
            {                                                                \
                const int32_t o_ = g0 + 16 * (8 * r_ + k_);                  \
                v_[k_] = (act && o_ >= 0) ? ld16(abase + o_) : make_uint4(0, 0, 0, 0); \
            }                                                                \
            _Pragma("unroll") for (int k_ = 0; k_ < 8; k_++)                 \
            {                                                                \
                uint32_t *w_ = win + 4 * (8 * r_ + k_);                      \
                w_[0] = v_[k_].x;                                            \
                w_[1] = v_[k_].y;                                            \
                w_[2] = v_[k_].z;                                            \
                w_[3] = v_[k_].w;                                            \
            }                                                                \
        }                                                                    \
        /* the bytes of the next refill into L2 meanwhile */                 \
        if (act && g0 >= 256)                                                \
        {                                                                    \
            zp_prefetch_l2(abase + g0 - 128);                                \
            zp_prefetch_l2(abase + g0 - 256);                                \
        }                                                                    \
        __syncwarp();                                                        \
    }
    ZP2B_FILL();
    /* the word holding the last byte: end mark, first bits */
    uint32_t hi = 0, lo = 0, cand = 0;
    int32_t  avail = 0, used = 0, total = 0;

    if (act)
    {
        uint32_t w = win[(npos - g0) >> 2];
        const uint32_t keep = delta + sn - (uint32_t) npos;     /* 1..4 valid low bytes */

        if (keep < 4)
            w &= (1u << (8u * keep)) - 1u;
        if (npos < delta)
            w &= ~0u << (8u * (uint32_t) (delta - npos));
        if ((w >> (8u * (keep - 1u))) == 0)
        {
            ok = false;                         /* no end mark in the last byte */
            act = false;
        }
        else
        {
            const int hbit = zs_highbit(w);

            hi = hbit ? w << (32 - hbit) : 0u;
            avail = hbit;
            total = (int32_t) ((sn - 1u) * 8u) + (hbit - 8 * (int) (keep - 1u));
            npos -= 4;
            cand = win[(npos - g0) >> 2];
        }
    }
    if (!act)
        cnt = 0;
    const uint16_t *huf = reinterpret_cast<const uint16_t *>(smem + ti * TB);
    const uint32_t sh = 32u - (uint32_t) tlog;

#ifdef CRYO_EMU
#define ZP_SHR_C(x, s) ((s) >= 32 ? 0u : (x) >> (s))
#define ZP_SHL_C(x, s) ((s) >= 32 ? 0u : (x) << (s))
#else
#define ZP_SHR_C(x, s) __funnelshift_rc((x), 0u, (uint32_t) (s))
#define ZP_SHL_C(x, s) __funnelshift_lc(0u, (x), (uint32_t) (s))
#endif
/* predicated refill: on = this lane still decodes */
#define ZP2B_REFILL(on)                                                      \
    {                                                                        \
        const bool p_ = (on) && avail <= 32;                                 \
        hi |= p_ ? ZP_SHR_C(cand, avail) : 0u;                               \
        lo = p_ ? ZP_SHL_C(cand, 32 - avail) : lo;                           \
        avail += p_ ? 32 : 0;                                                \
        npos -= p_ ? 4 : 0;                                                  \
        cand = win[(npos - g0) >> 2];                                        \
    }
#define ZP2B_DEC(sym, on)                                                    \
    {                                                                        \
        const uint32_t ent_ = huf[hi >> sh];                                 \
        const uint32_t nb_ = (on) ? ent_ >> 8 : 0u;                          \
        sym = ent_ & 0xFFu;                                                  \
        hi = __funnelshift_l(lo, hi, nb_);                                   \
        lo <<= nb_;                                                          \
        avail -= (int32_t) nb_;                                              \
        used += (int32_t) nb_;                                               \
    }
/* the window must hold the words of the next four symbols (at most two) and the fetch after them */
#define ZP2B_ENSURE_N(on, bytes)                                             \
    if (__any_sync(CRYO_FULL, (on) && npos - g0 < (bytes)))                  \
    {                                                                        \
        ZP2B_FILL();                                                         \
        cand = win[(npos - g0) >> 2];                                        \
    }
#define ZP2B_ENSURE(on) ZP2B_ENSURE_N(on, 8)
    {
        /* head: single symbols until dst + i is 4-byte aligned */
        const uint32_t head = act ? min((uint32_t) ((4u - ((uintptr_t) dst & 3u)) & 3u), cnt) : 0u;
        uint32_t i = 0;

        ZP2B_ENSURE(act);
#pragma unroll 1
        for (uint32_t k = 0; k < 3; k++)
        {
            const bool on = k < head;
            uint32_t   sy;

            ZP2B_REFILL(on);
            ZP2B_DEC(sy, on);
            if (on)
                dst[i++] = (uint8_t) sy;
        }
        /* quads, four symbols per 32-bit store; every lane runs the warp's longest count */
        const uint32_t quads = (cnt - i) >> 2;
        const uint32_t maxq = __reduce_max_sync(CRYO_FULL, quads);
        uint32_t *d4 = reinterpret_cast<uint32_t *>(dst + i);

#pragma unroll 1
        for (uint32_t q = 0; q < maxq; q++)
        {
            const bool on = q < quads;
            uint32_t   s0, s1, s2, s3;

            /* checked every fourth quad (the vote and its branch sit on every lane's chain): sixteen symbols take
             * 176 bits, six refills of a word each at most */
            if ((q & 3u) == 0)
                ZP2B_ENSURE_N(on, 32);
            ZP2B_REFILL(on);
            ZP2B_DEC(s0, on);
            ZP2B_DEC(s1, on);
            ZP2B_REFILL(on);
            ZP2B_DEC(s2, on);
            ZP2B_DEC(s3, on);
            if (on)
                d4[q] = s0 | (s1 << 8) | (s2 << 16) | (s3 << 24);
        }
        i += quads << 2;
        /* tail */
        const uint32_t tail = cnt - i;

        ZP2B_ENSURE(tail != 0);
#pragma unroll 1
        for (uint32_t k = 0; k < 3; k++)
        {
            const bool on = k < tail;
            uint32_t   sy;

            ZP2B_REFILL(on);
            ZP2B_DEC(sy, on);
            if (on)
                dst[i++] = (uint8_t) sy;
        }
    }
#undef ZP2B_FILL
#undef ZP2B_REFILL
#undef ZP2B_DEC
#undef ZP2B_ENSURE
#undef ZP2B_ENSURE_N
#undef ZP_SHR_C
#undef ZP_SHL_C
    /* every bit under the end mark consumed, no more, no less */
    if (act && used != total)
        ok = false;
    if (valid && !ok)
        a.flag[tf] = 1;
}

/* the low n bits of v, n = 0 .. 32 */
CRYO_DEV uint32_t zp_low_bits(uint32_t v, uint32_t n)
{
#ifdef CRYO_EMU
    return n >= 32u ? v : v & ((1u << n) - 1u);
#else
    uint32_t r;

    asm("bfe.u32 %0, %1, 0, %2;" : "=r"(r) : "r"(v), "r"(n));
    return r;
#endif
}

/* -------------------------------------------------------------- stage 3: sequences ---- */

/*
 * 3a: one CTA per (ZP_G frames x block index): the table descriptions of the blocks read one
 *     lane per block, then each table built by a warp in shared memory and copied to the
 *     block's slot of fsetab (global, LL u32[512] | OF u32[256] | ML u32[512]); logs and the
 *     bitstream offset go into the descriptor.
 * 3b: one LANE per block, 32 blocks (same block index of 32 frames) per warp, cells read
 *     through L1, (ll, ml, offset value) written to the frame's sequence area.
 */
#define ZP3_CELLS       1280u                   /* u32 cells per block slot */
#define ZP3A_WARPS      4u
#define ZP3A_CNT        388u                    /* per block: counts i16[64] x 3; odd word stride */
#define ZP3A_DESC       272u                    /* staged table descriptions per block */
#define ZP3A_OFF_DESC   (ZP_G * ZP3A_CNT)       /* 3 104: multiple of 16 */
#define ZP3A_OFF_CELLS  (ZP3A_OFF_DESC + ZP_G * ZP3A_DESC)      /* per warp: u32[512], the table being built */
#define ZP3A_OFF_WORK   (ZP3A_OFF_CELLS + ZP3A_WARPS * 2048u)   /* per warp: next u16[64] | cum u16[72] */
#define ZP3A_OFF_META   (ZP3A_OFF_WORK + ZP3A_WARPS * 272u)
#define ZP3A_SMEM       (ZP3A_OFF_META + ZP_G * 16u)

/*
 * stage 3a body: one CTA of ZP3A_WARPS warps, block index j of frames [g * ZP_G, g * ZP_G + ZP_G).
 * Phase A, warp 0: lane i reads the three table descriptions of frame g * ZP_G + i's block
 * (serial bit parsing: ZP_G blocks in lockstep).  Phase B: the warps share out the 3 * ZP_G tables
 * (three each; twelve each with 32 frames per CTA, 78 % of the stage in round 1);
 * each is built by a whole warp in shared memory and copied to the block's global slot.
 */
CRYO_DEV void zp_stage3a(const ZpArgs &a, uint32_t g, uint32_t j, uint8_t *smem, uint32_t tid)
{
    const uint32_t warp = tid >> 5, lane = tid & 31u;
    uint32_t *s_meta = reinterpret_cast<uint32_t *>(smem + ZP3A_OFF_META);      /* per block: info[3] | bitoff */

    if (warp == 0)
    {
        const uint32_t f = lane < ZP_G ? g * ZP_G + lane : a.n;
        uint32_t info[3] = {0, 0, 0}, bitoff = 0;   /* info: 1 << 31 | mode << 24 | rle sym or (nsym << 8 | log) */

        if (f < a.n && a.fr[(size_t) f * ZP_FF] > j)
        {
            const uint32_t *b = a.blk + ((size_t) f * ZP_MAXB + j) * ZP_BF;
            const uint32_t kind = b[ZPB_KIND];

            if ((kind & 3u) == 2u && b[ZPB_NSEQ] != 0)
            {
                /* the descriptions (three FSE headers: well under 256 bytes) -> shared memory; what
                 * follows them is the bitstream, which stage 3b reads */
                const uint8_t *gd = a.src + a.src_off[f] + b[ZPB_OFF] + b[ZPB_SEQOFF];
                const uint32_t full = b[ZPB_BSIZE] - b[ZPB_SEQOFF];
                const uint32_t dd = (uint32_t) ((uintptr_t) gd & 15u);
                uint8_t *sd = smem + ZP3A_OFF_DESC + lane * ZP3A_DESC;
                const uint32_t staged = full < ZP3A_DESC - 32u ? full : ZP3A_DESC - 32u;
                uint32_t left = staged;
                const uint8_t *p = sd + dd;
                bool bad = false;

                for (uint32_t k = 0; k < dd + staged; k += 16)
                    st16(sd + k, ld16(gd - dd + k));
#pragma unroll
                for (int t = 0; t < 3; t++)
                {
                    const uint32_t mode = (kind >> (14 - 2 * t)) & 3u;
                    const int max_log = t == 1 ? 8 : 9, max_sym = t == 0 ? 35 : t == 1 ? 31 : 52;

                    if (bad)
                        break;
                    if (mode == 0)
                        info[t] = 0x80000000u;
                    else if (mode == 1)
                    {
                        if (left < 1 || p[0] > max_sym)
                            bad = true;
                        else
                        {
                            info[t] = 0x80000000u | (1u << 24) | p[0];
                            p += 1;
                            left -= 1;
                        }
                    }
                    else
                    {
                        int32_t  nsym = 0, log = 0;
                        const uint32_t used = fse_read_counts(p, left, max_log, max_sym,
                                                              reinterpret_cast<int16_t *>(smem + lane * ZP3A_CNT + 128u * (uint32_t) t),
                                                              &nsym, &log);

                        if (used == 0)
                            bad = true;
                        else
                        {
                            info[t] = 0x80000000u | (2u << 24) | ((uint32_t) nsym << 8) | (uint32_t) log;
                            p += used;
                            left -= used;
                        }
                    }
                }
                if (bad)
                {
                    a.flag[f] = 1;
                    info[0] = 0;
                }
                bitoff = b[ZPB_SEQOFF] + (staged - left);
            }
        }
        if (lane < ZP_G)
        {
            s_meta[4 * lane + 0] = info[0];
            s_meta[4 * lane + 1] = info[1];
            s_meta[4 * lane + 2] = info[2];
            s_meta[4 * lane + 3] = bitoff;
        }
    }
    __syncthreads();
    uint32_t *cell = reinterpret_cast<uint32_t *>(smem + ZP3A_OFF_CELLS + warp * 2048u);
    uint16_t *next = reinterpret_cast<uint16_t *>(smem + ZP3A_OFF_WORK + warp * 272u);
    uint16_t *cum = next + 64;

    for (uint32_t w = warp; w < 3u * ZP_G; w += ZP3A_WARPS)
    {
        const uint32_t blkno = w / 3u, t = w % 3u;
        const uint32_t info = s_meta[4 * blkno + t];

        if (s_meta[4 * blkno] == 0)
            continue;
        const uint32_t mode = (info >> 24) & 3u;
        const uint32_t f = g * ZP_G + blkno;
        int32_t   logv = 0;

        __syncwarp();
        if (mode == 0)
        {
            const uint32_t n = t == 1 ? 32u : 64u, o = t == 0 ? 0u : t == 1 ? 64u : 96u;

            for (uint32_t k = lane; k < n; k += 32)
                cell[k] = a.predef[o + k];
            logv = t == 1 ? 5 : 6;
        }
        else if (mode == 1)
        {
            if (lane == 0)
                cell[0] = info & 0xFFu;
        }
        else
        {
            logv = (int32_t) (info & 0xFFu);
            fse_build_table_warp(cell, reinterpret_cast<const int16_t *>(smem + blkno * ZP3A_CNT + 128u * t),
                                 (int) ((info >> 8) & 0xFFu), logv, next, cum, lane);
        }
        __syncwarp();
        /* extra-bit count of every cell's code into bits 26..30 (as zsw_seq_table), and out to the slot */
        uint32_t *slot = a.fsetab + ((size_t) f * ZP_MAXB + j) * ZP3_CELLS + (t == 0 ? 0u : t == 1 ? 512u : 768u);

        for (uint32_t k = lane; k < (1u << logv); k += 32)
        {
            const uint32_t c = cell[k], sym = c & 0xFFu;
            const uint32_t xb = t == 1 ? sym : (t == 0 ? CRYO_GLD(ZS_LL_PACK[sym]) : CRYO_GLD(ZS_ML_PACK[sym])) >> 24;

            slot[k] = (c & 0x03FFFFFFu) | (xb << 26);
        }
        /* the logs: one word per block, assembled by whoever builds the block's tables (atomicOr: the
         * three tables of a block may be built by different warps) */
        if (lane == 0)
        {
            uint32_t *b = a.blk + ((size_t) f * ZP_MAXB + j) * ZP_BF;

            atomicOr(&b[ZPB_SLOGS], ((uint32_t) logv << (8 * t)) | 0x80000000u);
            if (t == 0)
                b[ZPB_BITOFF] = s_meta[4 * blkno + 3];
        }
    }
}

/*
 * stage 3b body: one warp, block index j of frames [g * LANES, g * LANES + LANES); lane i < LANES
 * walks frame g * LANES + i (8 blocks per warp).  The live cells of the three tables are packed LL | OF | ML into
 * CELLS u32 of shared memory per block.  The kernel is launched once per size class (small:
 * what libzstd emits for sparse blocks, 6/6/7-bit tables; large: the 9/8/9-bit maximum) and a
 * warp runs in the smallest class that holds all its blocks, so the small class keeps many
 * groups resident per SM.
 *
 * Like stage 2b the walk is written for lockstep: the bitstream of every lane is staged
 * through a 256-byte shared-memory window, refills of the 64-bit accumulator are predicated,
 * the next window word is fetched one refill ahead, and there is no branch or global load
 * inside a sequence (the 8-byte result store aside).
 */
#define ZP3B_SMALL      320u
#define ZP3B_LARGE      ZP3_CELLS
#define ZP3B_WIN        256u
#define ZP3B_WSTRIDE    (ZP3B_WIN / 4u + 1u)
#define ZP3B_SMEM(cells, lanes) ((lanes) * (cells) * 4u + (lanes) * ZP3B_WSTRIDE * 4u + 96u * 4u)   /* cells | windows | LL, ML code tables */
#define ZP3B_SMALL_LANES 8u                     /* blocks per warp in the small class (32 measured slower: 265 vs 221 us) */

template <uint32_t CELLS, uint32_t BELOW, uint32_t LANES>   /* groups of LANES frames needing > BELOW and <= CELLS cells */
CRYO_DEV void zp_stage3b(const ZpArgs &a, uint32_t g, uint32_t j, uint8_t *smem, uint32_t lane)
{
    const uint32_t f = g * LANES + lane;
    const uint32_t *b = nullptr;
    uint32_t logs = 0, nseq = 0, need = 0;
    bool     act = false;

    if (lane < LANES && f < a.n && a.fr[(size_t) f * ZP_FF] > j && a.flag[f] == 0)
    {
        b = a.blk + ((size_t) f * ZP_MAXB + j) * ZP_BF;
        nseq = b[ZPB_NSEQ];
        if ((b[ZPB_KIND] & 3u) == 2u && nseq != 0)
        {
            logs = b[ZPB_SLOGS];
            act = (logs & 0x80000000u) != 0;    /* else: the block's stage-3a warp flagged the frame */
            need = act ? (1u << (logs & 0xFFu)) + (1u << ((logs >> 8) & 0xFFu)) + (1u << ((logs >> 16) & 0x7Fu)) : 0u;
        }
    }
    /* the size class is decided for 32 consecutive frames (the small class's group), so that the
     * two launches agree on who takes what */
    uint32_t cneed = 0;

    {
        const uint32_t cf = (g * LANES / 32u) * 32u + lane;

        if (cf < a.n && a.fr[(size_t) cf * ZP_FF] > j && a.flag[cf] == 0)
        {
            const uint32_t *cb = a.blk + ((size_t) cf * ZP_MAXB + j) * ZP_BF;
            const uint32_t cl = cb[ZPB_SLOGS];

            if ((cb[ZPB_KIND] & 3u) == 2u && cb[ZPB_NSEQ] != 0 && (cl & 0x80000000u))
                cneed = (1u << (cl & 0xFFu)) + (1u << ((cl >> 8) & 0xFFu)) + (1u << ((cl >> 16) & 0x7Fu));
        }
    }
    const uint32_t gneed = __reduce_max_sync(CRYO_FULL, cneed);

    if (gneed == 0 || gneed > CELLS || gneed <= BELOW || __reduce_max_sync(CRYO_FULL, need) == 0)
        return;
    /* tables -> shared memory, the whole warp per block */
    for (uint32_t m = __ballot_sync(CRYO_FULL, act); m; m &= m - 1)
    {
        const int      i = __ffs((int) m) - 1;
        const uint32_t li = __shfl_sync(CRYO_FULL, logs, i);
        const uint32_t nl = 1u << (li & 0xFFu), no = 1u << ((li >> 8) & 0xFFu), nm = 1u << ((li >> 16) & 0x7Fu);
        const uint32_t *slot = a.fsetab + ((size_t) (g * LANES + (uint32_t) i) * ZP_MAXB + j) * ZP3_CELLS;
        uint32_t *cells = reinterpret_cast<uint32_t *>(smem) + (uint32_t) i * CELLS;

        /* four loads in flight per lane (one at a time, a block's tables cost ten L2 round trips: a fifth of the
         * stage's stall samples on the sparse table, profiles/r02n) */
        for (uint32_t base = 0; base < nl + no + nm; base += 128u)
        {
            uint32_t v[4];

#pragma unroll
            for (uint32_t q = 0; q < 4; q++)
            {
                const uint32_t k = base + 32u * q + lane;

                v[q] = k >= nl + no + nm ? 0u : slot[k < nl ? k : k < nl + no ? 512u + (k - nl) : 768u + (k - nl - no)];
            }
#pragma unroll
            for (uint32_t q = 0; q < 4; q++)
            {
                const uint32_t k = base + 32u * q + lane;

                if (k < nl + no + nm)
                    cells[k] = v[q];
            }
        }
    }
    const uint32_t ll_log = logs & 0xFFu, of_log = (logs >> 8) & 0xFFu, ml_log = (logs >> 16) & 0x7Fu;
    const uint32_t *llt = reinterpret_cast<const uint32_t *>(smem) + (lane & (LANES - 1u)) * CELLS;
    const uint32_t *oft = llt + (1u << ll_log), *mlt = oft + (1u << of_log);
    /* this lane's bitstream src[0, sn); positions are byte offsets from abase = src rounded down to 16 */
    const uint8_t *src = nullptr;
    uint32_t  sn = 0;
    bool      bad = false;

    if (act)
    {
        const uint32_t bitoff = b[ZPB_BITOFF];

        src = a.src + a.src_off[f] + b[ZPB_OFF] + bitoff;
        sn = b[ZPB_BSIZE] - bitoff;
        if (sn == 0)
        {
            bad = true;
            act = false;
        }
    }
    /*
     * The walk.  What is serial in a sequence bitstream is short: the three states name the cells, the
     * cells say how many bits the sequence takes, and the next states are three fields of the LAST bits
     * it takes.  The values of the extra bits (offset, lengths) are not needed to get there.  So the
     * cursor is a bit position P (the stream is read from its end: bits at and above P are consumed),
     * any field is fetched by position from the lane's window in shared memory -- two words and a funnel
     * shift -- and only  cells -> bit counts -> position of the state field -> next states  is a
     * dependent chain (about a hundred cycles); the fetches of the extra bits, the baselines and the
     * store hang off it.  Round 1 drew every field from one shift-register accumulator with three
     * conditional refills per sequence: 154 dependent instructions, 1 300 cycles per sequence.
     */
    /* a window per working lane (the others only read: with windows for all 32 the small class took 18.9 KB a warp, and
     * the sequence warps of the sparse table left room for two of its three literal warps per SM) */
    uint32_t *win = reinterpret_cast<uint32_t *>(smem + LANES * CELLS * 4u) + (lane & (LANES - 1u)) * ZP3B_WSTRIDE;
    const uint8_t *abase = src - ((uintptr_t) src & 15u);
    const int32_t  delta = (int32_t) ((uintptr_t) src & 15u);
    int32_t  P = 0;                             /* bit offset from abase */
    int32_t  g0 = 0;                            /* the window holds bytes [g0, g0 + ZP3B_WIN) from abase; multiple of 16 */

    if (act)
    {
        const uint32_t lastb = src[sn - 1u];

        if (lastb == 0)
        {
            bad = true;                         /* no end mark in the last byte */
            act = false;
        }
        else
            P = 8 * (delta + (int32_t) sn - 1) + zs_highbit(lastb);
    }
    if (!act)
        nseq = 0;
    const int32_t Pend = 8 * delta;             /* a valid walk ends exactly here */

/* the window, so that the cursor's byte is among its last sixteen (bytes in front of abase read as zero) */
#define ZP3B_FILL()                                                          \
    {                                                                        \
        __syncwarp();                                                        \
        g0 = ((P >> 3) & ~15) - (int32_t) (ZP3B_WIN - 16u);                  \
        _Pragma("unroll") for (int r_ = 0; r_ < 2; r_++)                     \
        {                                                                    \
            uint4 v_[8];                                                     \
            _Pragma("unroll") for (int k_ = 0; k_ < 8; k_++)                 \
            {                                                                \
                const int32_t o_ = g0 + 16 * (8 * r_ + k_);                  \
                v_[k_] = (act && o_ >= 0) ? ld16(abase + o_) : make_uint4(0, 0, 0, 0); \
            }                                                                \
            if (lane < LANES)                                                \
                _Pragma("unroll") for (int k_ = 0; k_ < 8; k_++)             \
                {                                                            \
                    uint32_t *w_ = win + 4 * (8 * r_ + k_);                  \
                    w_[0] = v_[k_].x;                                        \
                    w_[1] = v_[k_].y;                                        \
                    w_[2] = v_[k_].z;                                        \
                    w_[3] = v_[k_].w;                                        \
                }                                                            \
        }                                                                    \
        /* the bytes of the next refill into L2 meanwhile */                 \
        if (act && g0 >= 256)                                                \
        {                                                                    \
            zp_prefetch_l2(abase + g0 - 128);                                \
            zp_prefetch_l2(abase + g0 - 256);                                \
        }                                                                    \
        __syncwarp();                                                        \
    }
/* n <= 32 bits whose lowest is bit q of the stream (n == 0: nothing) */
#define ZP3B_BITS(dst, q, n)                                                 \
    {                                                                        \
        const int32_t   q_ = (q);                                            \
        const uint32_t *w_ = win + ((q_ >> 5) - (g0 >> 2));                  \
        dst = zp_low_bits(__funnelshift_r(w_[0], w_[1], (uint32_t) q_ & 31u), (n)); \
    }
/* four sequences take 4 x 85 bits at most: 43 bytes below the cursor */
#define ZP3B_ENSURE(on)                                                      \
    {                                                                        \
        if (act && P < Pend)                                                 \
        {                                                                    \
            bad = true;                         /* read past the start of the stream */ \
            nseq = 0;                                                        \
        }                                                                    \
        if (__any_sync(CRYO_FULL, (on) && (P >> 3) - g0 < 48))               \
            ZP3B_FILL();                                                     \
    }
    ZP3B_FILL();
    /* code -> baseline | extra bits << 24, from shared memory in the loop */
    uint32_t *packs = reinterpret_cast<uint32_t *>(smem + LANES * CELLS * 4u + LANES * ZP3B_WSTRIDE * 4u);

    for (uint32_t k = lane; k < 36u + 53u; k += 32)
        packs[k] = k < 36u ? CRYO_GLD(ZS_LL_PACK[k]) : CRYO_GLD(ZS_ML_PACK[k - 36u]);
    __syncwarp();
    uint32_t sl = 0, so = 0, sm = 0, mlsum = 0;

    {
        /* the initial states: LL, OF, ML (26 bits at most) */
        const uint32_t n0 = act ? ll_log + of_log + ml_log : 0u;
        uint32_t t;

        P -= (int32_t) n0;
        ZP3B_BITS(t, P, n0);
        sl = act ? t >> (of_log + ml_log) : 0u;
        so = act ? (t >> ml_log) & ((1u << of_log) - 1u) : 0u;
        sm = act ? t & ((1u << ml_log) - 1u) : 0u;
    }
    uint64_t *out = act ? a.seq + a.seqbase[f] + b[ZPB_SEQPOS] : nullptr;
    const uint32_t maxseq = __reduce_max_sync(CRYO_FULL, nseq);

#pragma unroll 1
    for (uint32_t i = 0; i < maxseq; i++)
    {
        if ((i & 3u) == 0)
            ZP3B_ENSURE(i < nseq);
        const bool on = i < nseq;
        const bool more = i + 1 < nseq;
        const uint32_t cl = llt[sl], co = oft[so], cm = mlt[sm];
        const uint32_t xo = co >> 26, xm = cm >> 26, xl = cl >> 26;
        const uint32_t nbl = (cl >> 8) & 0xFFu, nbm = (cm >> 8) & 0xFFu, nbo = (co >> 8) & 0xFFu;
        /* offset extra bits (27 at most: beyond any window ZSTD_decompress accepts), then match-length and
         * literal-length extra bits (16 + 16 at most), then the state updates LL, ML, OF (9 + 9 + 8 at most) */
        const uint32_t n1 = on ? (xo > 27 ? 27u : xo) : 0u, n2 = on ? xm + xl : 0u, n3 = more ? nbl + nbm + nbo : 0u;
        const int32_t  Pa = P - (int32_t) n1, Pb = Pa - (int32_t) n2, Pc = Pb - (int32_t) n3;
        uint32_t t3, ov, t2;

        ZP3B_BITS(t3, Pc, n3);
        sl = more ? ((cl >> 16) & 0x3FFu) + (t3 >> (nbm + nbo)) : sl;
        sm = more ? ((cm >> 16) & 0x3FFu) + ((t3 >> nbo) & ((1u << nbm) - 1u)) : sm;
        so = more ? ((co >> 16) & 0x3FFu) + (t3 & ((1u << nbo) - 1u)) : so;
        P = Pc;
        /* off the chain: the values */
        if (on && xo > 27)
            bad = true;
        ZP3B_BITS(ov, Pa, n1);
        ZP3B_BITS(t2, Pb, n2);
        ov += 1u << (xo & 31u);
        /* (a lane without work reads stale cells: keep its table indices in range) */
        const uint32_t pl = packs[on ? cl & 0xFFu : 0u], pm = packs[36u + (on ? cm & 0xFFu : 0u)];
        const uint32_t ml = (pm & 0xFFFFFFu) + (t2 >> xl);
        const uint32_t ll = (pl & 0xFFFFFFu) + (t2 & ((1u << xl) - 1u));

        if (on)
        {
            out[i] = (uint64_t) ll | ((uint64_t) ml << 17) | ((uint64_t) (ov & 0x1FFFFFFFu) << 35);
            mlsum += ml;
        }
    }
    if (act)
        a.blk[((size_t) f * ZP_MAXB + j) * ZP_BF + ZPB_OUTSZ] += mlsum;     /* regen + match bytes (< 2^32) */
#undef ZP3B_FILL
#undef ZP3B_BITS
#undef ZP3B_ENSURE
    /* every bit under the end mark consumed, no more, no less */
    if (act && nseq != 0 && P != Pend)
        bad = true;
    if (bad)
        a.flag[f] = 1;
}

/* ---------------------------------------------------------------- stage 4: execute ---- */

#ifndef ZP4_WARPS
#define ZP4_WARPS       8u
#endif
#define ZP4_THREADS     (32u * ZP4_WARPS)
#define ZP4_LITWIN      1024u                   /* literal window (the slow path uses its first ZSW_LITWIN bytes) */
#define ZP4_DESC        (ZP_MAXB * ZP_BF * 4u + ZP_JOBS * 16u)  /* the frame's block descriptors | the runs handed to stage 0 (position, length, byte) */
#define ZP4_PER_WARP    (WX_RING + ZP4_LITWIN + ZP4_DESC)
#define ZP4_SMEM        (ZP4_WARPS * ZP4_PER_WARP)
#define ZP4_SPAN        1024u                   /* output bytes of one sub-batch: it is written ahead of o.pos in the ring */
#define ZP4_BIG_LL      96u                     /* longer runs leave the batch path */
#define ZP4_BIG_ML      256u
#ifdef CRYO_EMU
#define ZP4_SPINS        4u
#else
#define ZP4_SPINS        8192u                  /* x 256 ns: about 2 ms, then stage 4 writes the blocks itself */
#endif
#define ZP4_LANE_ML     48u                     /* independent matches up to this long are copied by one lane each */

/*
 * Stage 4 reads what stage 0 writes only through match sources below `guard` (the end of the last
 * block it skipped).  Before the first such read it waits until stage 0 has published the frame's
 * blocks, or writes them itself if that takes too long (idempotent).  A frame whose matches never
 * reach back into a skipped block never waits: the zero run that follows the RLE blocks of a
 * sparse cryo block is recognised as a fill of the RLE byte (see rle_lo / rle_hi below).
 */
/* stage 0 owes this frame something stage 4 has not seen published: blocks (skipped) or jobs (njobs) */
#define ZP4_OWED() (skipped > confirmed || njobs > jconfirmed)
#define ZP4_CONFIRM()                                                                        \
    if (ZP4_OWED())                                                                          \
    {                                                                                        \
        uint32_t ok_ = 0;                                                                    \
                                                                                             \
        if (lane == 0)                                                                       \
            for (uint32_t spin_ = 0; spin_ < ZP4_SPINS; spin_++)                             \
            {                                                                                \
                ok_ = zp_ld_acquire(a.pf_done + f) >= skipped;                               \
                for (uint32_t kk_ = jconfirmed; kk_ < njobs && ok_; kk_++)                   \
                    ok_ = (zp_ld_acquire(a.jobs + 4u * (size_t) sjobs[4u * kk_ + 3u] + 3) & ZP_JOB_DONE) != 0; \
                if (ok_)                                                                     \
                    break;                                                                   \
                __nanosleep(256);                                                            \
            }                                                                                \
        ok_ = __shfl_sync(CRYO_FULL, ok_, 0);                                                \
        __threadfence();        /* every lane's later reads of stage 0's output after lane 0's acquire */ \
        if (!ok_)                                                                            \
        {                                                                                    \
            uint32_t at_ = 0;                                                                \
                                                                                             \
            for (uint32_t jj_ = 0; jj_ < j; jj_++)                                           \
            {                                                                                \
                const uint32_t *bb_ = a.blk + ((size_t) f * ZP_MAXB + jj_) * ZP_BF;          \
                                                                                             \
                if ((skipmask >> jj_) & 1u)                                                  \
                {                                                                            \
                    if ((bb_[ZPB_KIND] & 3u) == 0)                                           \
                        team_copy(o.out + at_, fin + bb_[ZPB_OFF], bb_[ZPB_BSIZE], lane, 32); \
                    else                                                                     \
                        team_fill_byte(o.out + at_, fin[bb_[ZPB_OFF]], bb_[ZPB_BSIZE], lane, 32); \
                }                                                                            \
                at_ += bb_[ZPB_OUTSZ];                                                       \
            }                                                                                \
            for (uint32_t kk_ = 0; kk_ < njobs; kk_++)                                       \
                team_fill_byte(o.out + sjobs[4u * kk_], (uint8_t) sjobs[4u * kk_ + 2u], sjobs[4u * kk_ + 1u], lane, 32); \
            __syncwarp();                                                                    \
        }                                                                                    \
        confirmed = skipped;                                                                 \
        jconfirmed = njobs;                                                                  \
        hull_lo = ~0u;                                                                       \
    }

/* stage 4 body: one warp, frame f */
CRYO_DEV void zp_stage4(const ZpArgs &a, uint32_t f, uint8_t *smem, uint32_t lane)
{
    if (f >= a.n)
        return;
    /* (loaded side by side: a chain of || would make them four round trips) */
    const int32_t  method = a.methods[f];
    const uint32_t flagged = a.flag[f], routed = a.fr[(size_t) f * ZP_FF + 3];
    const uint32_t nb = a.fr[(size_t) f * ZP_FF], cap = a.cap;

    if (method != ZP_METHOD_ZSTD || flagged != 0 || routed != 0)
        return;
    const uint8_t *in = a.src + a.src_off[f], *fin = in;    /* fin: for ZP4_CONFIRM, where `in` is shadowed */
    WOut     o;
    int      err = ST_OK;
    uint32_t rep0 = 1, rep1 = 4, rep2 = 8;
    uint32_t skipped = 0, confirmed = 0, skipmask = 0;     /* blocks left to stage 0 */
    uint32_t guard = 0;                                    /* positions below may not be written yet (see ZP4_CONFIRM) */
    uint32_t hull_lo = ~0u;                                /* ... and none below this is stage 0's to write: [hull_lo, guard) spans what it owes */
    uint32_t rle_lo = 0, rle_hi = 0;                       /* [rle_lo, rle_hi): skipped RLE blocks, all bytes = rle_byte */
    uint8_t  rle_byte = 0;

    const uint32_t le_mask = 0xFFFFFFFFu >> (31u - lane);  /* lanes up to this one */
    /*
     * The frame's block descriptors -> shared memory in one go.  Read from global memory block by block, field
     * after field, they were two dozen dependent round trips per frame while the raw / RLE stage keeps HBM busy:
     * a tenth of the executor's stall samples (profiles/r02q).
     */
    uint32_t *sdesc = reinterpret_cast<uint32_t *>(smem + WX_RING + ZP4_LITWIN);
    uint32_t *sjobs = sdesc + ZP_MAXB * ZP_BF;
    uint32_t  njobs = 0, jconfirmed = 0;                   /* runs handed to stage 0; how many of them it is known to have written */

    {
        const uint32_t *gdesc = a.blk + (size_t) f * ZP_MAXB * ZP_BF;
        const uint32_t words = nb * ZP_BF;

        for (uint32_t k0 = 0; k0 < words; k0 += 128u)
        {
            uint32_t v[4];

#pragma unroll
            for (uint32_t q = 0; q < 4; q++)
                v[q] = k0 + 32u * q + lane < words ? gdesc[k0 + 32u * q + lane] : 0u;
#pragma unroll
            for (uint32_t q = 0; q < 4; q++)
                if (k0 + 32u * q + lane < words)
                    sdesc[k0 + 32u * q + lane] = v[q];
        }
        __syncwarp();
    }
    wx_init(o, a.dst + (size_t) f * a.dst_stride, cap, smem);
    for (uint32_t j = 0; j < nb && err == ST_OK; j++)
    {
        const uint32_t *b = sdesc + j * ZP_BF;
        const uint32_t off = b[ZPB_OFF], bsize = b[ZPB_BSIZE], kind = b[ZPB_KIND], type = kind & 3u;

        if (type < 2)
        {
            if (bsize > cap - o.pos)
            {
                err = ST_OUTPUT;
                break;
            }
            if (bsize == 0)
                continue;
            if (b[ZPB_SPECPOS] != ~0u)
            {
                /* stage 0 writes this block here (it may not have yet): move on without touching
                 * the output, the ring's tail comes from the block's own description */
                /* written before this kernel started (the early pass of stage 0 guessed its position right):
                 * nothing to wait for, whoever reads it */
                const bool early = zp_early_frame(a, f) && b[ZPB_SPECAT] == o.pos;

                wx_drain_all(o, lane);
                o.pos += bsize;
                o.flushed = o.pos & ~15u;
                o.lo = o.flushed;
                if (lane < o.pos - o.flushed)
                    o.ring[(o.flushed + lane) & WX_RMASK] = type == 0 ? in[off + bsize - (o.pos - o.flushed) + lane]
                                                                      : (uint8_t) b[ZPB_RLEBYTE];
                __syncwarp();
                if (!early)
                {
                    if (!ZP4_OWED())
                        hull_lo = o.pos - bsize;
                    skipped++;
                    skipmask |= 1u << j;
                    guard = o.pos;
                }
                /* consecutive RLE blocks of one byte form one range whose content is known */
                if (type == 1 && rle_hi == o.pos - bsize && rle_byte == (uint8_t) b[ZPB_RLEBYTE] && rle_hi > rle_lo)
                    rle_hi = o.pos;
                else if (type == 1)
                {
                    rle_lo = o.pos - bsize;
                    rle_hi = o.pos;
                    rle_byte = (uint8_t) b[ZPB_RLEBYTE];
                }
                else
                    rle_lo = rle_hi = 0;
                continue;
            }
            if (type == 0)
                wx_literals(o, in + off, bsize, lane);
            else
                wx_fill_byte(o, (uint8_t) b[ZPB_RLEBYTE], bsize, lane);
            continue;
        }
        const uint32_t lt = (kind >> 2) & 3u, regen = b[ZPB_REGEN], nseq = b[ZPB_NSEQ];
        const uint32_t block_start = o.pos;
        const uint8_t *lit_base = lt >= 2 ? a.lit + (size_t) f * a.lit_stride + b[ZPB_LITPOS]
                                          : in + off + b[ZPB_LHDR];
        ZswLits  L;

        L.rle = lt == 1;
        L.rle_byte = lt == 1 ? lit_base[0] : (uint8_t) 0;
        L.abase = lit_base - ((uintptr_t) lit_base & 15u);
        L.delta = (uint32_t) ((uintptr_t) lit_base & 15u);
        L.n = regen;
        L.pos = 0;
        L.win = smem + WX_RING;
        L.wvalid = false;
        L.wbase = 0;
        L.lim = (L.delta + regen + 15u) & ~15u;

        if (nseq)
        {
            const uint64_t *sq = a.seq + a.seqbase[f] + b[ZPB_SEQPOS];
            uint64_t nxt = lane < nseq ? sq[lane] : 0ull;
            uint32_t lpos = 0;

            /*
             * 32 sequences per load, lane k = sequence done + k.  Offsets (repeat codes), output
             * and literal positions and the format checks are computed for the whole load first
             * (a serial pass over the offset codes, two warp scans, one vote).  Then runs of
             * ordinary sequences are executed as sub-batches written ahead of o.pos in the ring:
             * every lane copies the literals of its own sequence (they come from one contiguous
             * piece of the literal buffer, staged in the window), then the matches follow in
             * order, each a warp-wide move.  Long runs, and anything that does not fit the ring,
             * take the general executor one sequence at a time.
             */
            for (uint32_t done = 0; done < nseq && err == ST_OK; done += 32)
            {
                const uint32_t g = nseq - done < 32u ? nseq - done : 32u;
                const uint64_t cur = nxt;

                if (done + 32u + lane < nseq)
                    nxt = sq[done + 32u + lane];
#ifndef CRYO_EMU
                /* HBM is busy with the raw / RLE stage's stores while this runs, and a load that has to
                 * go there is slow: ask L2 for the sequences and the literals a few loads ahead (one
                 * 128-byte line per lane) */
                if ((a.pf_hint & 2u) && done + 64u + 16u * lane < nseq)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(sq + done + 64u + 16u * lane));
                if ((a.pf_hint & 2u) && lt != 1 && lpos + 1024u + 128u * lane < regen)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(lit_base + lpos + 1024u + 128u * lane));
#endif
                const bool     have = lane < g;
                const uint32_t my_ll = have ? (uint32_t) cur & 0x1FFFFu : 0u;
                const uint32_t my_ml = have ? (uint32_t) (cur >> 17) & 0x3FFFFu : 0u;
                const uint32_t my_ov = (uint32_t) (cur >> 35);
                uint32_t my_off = 0;

                /* repeat offsets, in order (RFC 8878 3.1.1.5) */
                {
                    const uint32_t zmask = __ballot_sync(CRYO_FULL, my_ll == 0);
                    /* the sequences in front of the load's first repeat code only push their offsets: done at once */
                    const uint32_t repmask = __ballot_sync(CRYO_FULL, have && my_ov <= 3u);
                    const uint32_t kfirst = repmask ? (uint32_t) __ffs((int) repmask) - 1u : g;

                    if (kfirst)
                    {
                        const uint32_t o1 = __shfl_sync(CRYO_FULL, my_ov, (int) (kfirst - 1u)) - 3u;
                        const uint32_t o2 = __shfl_sync(CRYO_FULL, my_ov, (int) ((kfirst - 2u) & 31u)) - 3u;
                        const uint32_t o3 = __shfl_sync(CRYO_FULL, my_ov, (int) ((kfirst - 3u) & 31u)) - 3u;

                        if (lane < kfirst)
                            my_off = my_ov - 3u;
                        if (kfirst >= 3u)
                        {
                            rep2 = o3;
                            rep1 = o2;
                        }
                        else if (kfirst == 2u)
                        {
                            rep2 = rep0;
                            rep1 = o2;
                        }
                        else
                        {
                            rep2 = rep1;
                            rep1 = rep0;
                        }
                        rep0 = o1;
                    }
                    for (uint32_t k = kfirst; k < g; k++)
                    {
                        const uint32_t ov = __shfl_sync(CRYO_FULL, my_ov, (int) k);
                        uint32_t moff;

                        if (ov > 3)
                        {
                            moff = ov - 3;
                            rep2 = rep1;
                            rep1 = rep0;
                            rep0 = moff;
                        }
                        else
                        {
                            const uint32_t idx = ov - 1 + ((zmask >> k) & 1u);

                            if (idx == 0)
                                moff = rep0;
                            else
                            {
                                moff = idx == 1 ? rep1 : idx == 2 ? rep2 : rep0 - 1;
                                if (idx > 1)
                                    rep2 = rep1;
                                rep1 = rep0;
                                rep0 = moff;
                            }
                        }
                        if (lane == k)
                            my_off = moff;
                    }
                }
                /* positions: inclusive scans of ll + ml and of ll */
                uint32_t cum = my_ll + my_ml, lcum = my_ll;

#pragma unroll
                for (uint32_t d = 1; d < 32; d <<= 1)
                {
                    const uint32_t c1 = __shfl_up_sync(CRYO_FULL, cum, d), c2 = __shfl_up_sync(CRYO_FULL, lcum, d);

                    if (lane >= d)
                    {
                        cum += c1;
                        lcum += c2;
                    }
                }
                const uint32_t base_pos = o.pos, base_lit = lpos;
                const uint32_t my_start = base_pos + cum - my_ll - my_ml, my_mpos = my_start + my_ll;
                const uint32_t my_epos = base_pos + cum;                /* < 2^28: no wrap */
                const uint32_t my_lit = base_lit + lcum - my_ll;        /* first literal of this sequence */

                if (__any_sync(CRYO_FULL, have && ((base_lit + lcum > regen) | (my_epos > cap) |
                                                   (my_epos - block_start > ZS_MAXBLOCK) |
                                                   (my_off - 1u >= my_mpos))))
                {
                    err = ST_FORMAT;            /* which rule, and the status, are the fallback's to say */
                    break;
                }
                const uint32_t bigmask = __ballot_sync(CRYO_FULL, !have || my_ll > ZP4_BIG_LL || my_ml > ZP4_BIG_ML);
                uint32_t k0 = 0;
                bool     need_confirm = false;

                while (k0 < g)
                {
                    if (need_confirm)
                    {
                        ZP4_CONFIRM();          /* the only expansion: the sites below come back here */
                        need_confirm = false;
                    }
                    /* the longest run of ordinary sequences from k0 whose output fits the span and
                     * whose literals fit the window */
                    const uint32_t pos0 = __shfl_sync(CRYO_FULL, my_start, (int) k0);
                    const uint32_t lit0 = __shfl_sync(CRYO_FULL, my_lit, (int) k0);
                    const uint32_t lip0 = L.delta + lit0, wb = lip0 & ~15u;
                    const bool     fits = my_epos - pos0 <= ZP4_SPAN && L.delta + my_lit + my_ll - wb <= ZP4_LITWIN;
                    const uint32_t stop = (bigmask | __ballot_sync(CRYO_FULL, !fits)) & (~0u << k0);
                    const uint32_t k1 = stop ? (uint32_t) __ffs((int) stop) - 1u : 32u;   /* first lane not in the run */

                    if (k1 == k0 || L.rle || pos0 != o.pos)
                    {
                        /* one sequence through the general executor */
                        const uint32_t ll = __shfl_sync(CRYO_FULL, my_ll, (int) k0);
                        const uint32_t ml = __shfl_sync(CRYO_FULL, my_ml, (int) k0);
                        const uint32_t moff = __shfl_sync(CRYO_FULL, my_off, (int) k0);

                        const uint32_t sp1 = o.pos + ll - moff, need = ml < moff ? ml : moff;
                        const bool     known = sp1 >= rle_lo && sp1 + need <= rle_hi;
                        /* a copy that starts inside the known range and runs out of its end (the zeros in front of a block's
                         * first tuple and the tuple's first bytes): the known part is a fill, the rest an ordinary match */
                        const uint32_t head_known = (!known && moff >= ml && sp1 >= rle_lo && sp1 < rle_hi) ? rle_hi - sp1 : 0u;

                        /* (a source the ring still holds is read from there whether or not stage 0 has written it) */
                        if (!known && sp1 + head_known < guard && sp1 + need > hull_lo && ZP4_OWED() &&
                            !(ll < WX_BULK && ml < WX_BULK && moff <= WX_RING - 64u && sp1 >= o.lo))
                        {
                            need_confirm = true;
                            continue;
                        }
                        L.pos = lit0;
                        zsw_lits_emit(o, L, ll, lane);
                        if (known)
                            wx_fill_byte(o, rle_byte, ml, lane);        /* a copy of known bytes: no read */
                        else if (a.jobs && moff == 1u && ml >= ZP_JOB_MIN && njobs < ZP_JOBS && o.pos > o.lo &&
                                 zp_ld_acquire(reinterpret_cast<const uint32_t *>(a.seq_alloc) + ZPC_SERVERS) != 0)
                        {
                            /* a long run of the previous byte: stage 0's job (see ZP_JOBS) */
                            const uint8_t  rb = o.ring[(o.pos - 1u) & WX_RMASK];
                            const uint32_t at = o.pos;

                            wx_drain_all(o, lane);
                            if (lane == 0)
                            {
                                uint32_t *ctl = reinterpret_cast<uint32_t *>(a.seq_alloc);
                                const uint32_t slot = atomicAdd(ctl + ZPC_JOB_TAIL, 1u);
                                uint32_t *job = a.jobs + 4u * (size_t) slot;

                                sjobs[4u * njobs + 3u] = slot;
                                job[0] = f;
                                job[1] = at;
                                job[2] = ml;
                                __threadfence();
                                zp_st_release(job + 3, (uint32_t) rb | ZP_JOB_READY);
                                sjobs[4u * njobs] = at;
                                sjobs[4u * njobs + 1u] = ml;
                                sjobs[4u * njobs + 2u] = rb;
                            }
                            /* the known range goes on through a few literals of the same byte (a sparse block's second
                             * zero run starts one literal zero after its RLE blocks) */
                            const uint32_t gap = at - rle_hi;
                            const bool     joins = rle_hi > rle_lo && rle_byte == rb && at >= rle_hi && gap <= 8u && rle_hi >= o.lo &&
                                                   __all_sync(CRYO_FULL, lane >= gap || o.ring[(rle_hi + lane) & WX_RMASK] == rb);

                            wx_after_fill(o, ml, rb, lane);
                            if (!ZP4_OWED())
                                hull_lo = at;
                            njobs++;
                            guard = o.pos;
                            if (joins)
                                rle_hi = o.pos;
                            else
                            {
                                rle_lo = at;
                                rle_hi = o.pos;
                                rle_byte = rb;
                            }
                        }
                        else if (head_known)
                        {
                            wx_fill_byte(o, rle_byte, head_known, lane);
                            wx_match(o, moff, ml - head_known, lane);
                        }
                        else
                            wx_match(o, moff, ml, lane);
                        k0++;
                        continue;
                    }
                    const bool     in = lane >= k0 && lane < k1;

                    const uint32_t end = __shfl_sync(CRYO_FULL, my_epos, (int) (k1 - 1u));
                    const uint32_t floor0 = end + 64u > WX_RING ? end + 64u - WX_RING : 0u;
                    const uint32_t floor = floor0 > o.lo ? floor0 : o.lo;

                    /* stage 0's bytes are needed by a source that is below what it may not have written yet (guard), older
                     * than the ring (floor) and not inside the range whose content is known (those bytes are not read) */
                    {
                        const uint32_t s0 = my_mpos - my_off, sl = my_ml < my_off ? my_ml : my_off;
                        /* (the bytes of a source that lie in the known range are not read: what counts is where the rest begins) */
                        const uint32_t u0 = s0 >= rle_lo && s0 < rle_hi ? rle_hi : s0;

                        if (ZP4_OWED() &&
                            __any_sync(CRYO_FULL, in && u0 < s0 + sl && u0 < guard && s0 + sl > hull_lo && u0 < floor))
                        {
                            need_confirm = true;
                            continue;
                        }
                    }
                    const uint32_t litend = __shfl_sync(CRYO_FULL, my_lit + my_ll, (int) (k1 - 1u));

                    /* literals of the run -> window (coalesced), then every lane places its own */
                    __syncwarp();
                    for (uint32_t w = 16u * lane; wb + w < L.delta + litend; w += 512u)
                        if (wb + w < L.lim)
                            st16(L.win + w, ld16(L.abase + wb + w));
                    L.wvalid = false;           /* the slow path's window bookkeeping no longer holds */
                    __syncwarp();
                    /*
                     * The run byte by byte, 32 consecutive output bytes per step, one per lane (round 1
                     * gave every lane its own sequence and moved its bytes one after the other: the
                     * lanes with short sequences idled, and the moves were most of the executor's 84 K
                     * instructions per sparse frame).  Which sequence a byte belongs to: the ends of the
                     * sequences that fall into the step's 32 bytes are marked in a word (one warp OR),
                     * and a lane counts the marks at or below its own byte.  A literal byte comes from
                     * the window, a match byte from its distance back: the ring holds everything before
                     * the step (the steps run in order), older bytes are in global memory, and a source
                     * INSIDE the step is another lane's byte -- those are resolved by pointer jumping
                     * over the lanes (a chain of overlapping copies of any length takes five rounds).
                     */
                    const uint32_t span = end - pos0;
                    const uint32_t e_rel = my_epos - pos0;
                    /* match start | (window index of the first literal - start, biased) of this lane's sequence */
                    const uint32_t my_pack = ((my_mpos - pos0) & 0xFFFFu) |
                                             ((L.delta + my_lit - wb + 2048u - (my_start - pos0)) << 16);
                    /* ring and window are one array (the window follows the ring): a byte's source is one index */
                    const bool     far = __any_sync(CRYO_FULL, in && my_mpos - my_off < floor);    /* sources in global memory */
                    uint32_t       seqbase = k0;

                    for (uint32_t R = 0; R < span; R += 32)
                    {
                        const uint32_t d = e_rel - R;
                        const uint32_t marks = __reduce_or_sync(CRYO_FULL, (in && d < 32u) ? 1u << d : 0u);
                        const uint32_t idx = (seqbase + __popc(marks & le_mask)) & 31u;
                        const uint32_t s_pack = __shfl_sync(CRYO_FULL, my_pack, (int) idx);
                        const uint32_t s_off = __shfl_sync(CRYO_FULL, my_off, (int) idx);
                        const uint32_t r = R + lane, p = pos0 + r, x = p - s_off;
                        const bool     live = r < span, lit = r < (s_pack & 0xFFFFu);
                        const bool     inside = live && !lit && s_off <= lane;     /* source in this step */
                        const uint32_t at = lit ? r + (s_pack >> 16) + (WX_RING - 2048u) : x & WX_RMASK;
                        uint32_t       v = 0;

                        seqbase += __popc(marks);
                        if (live && !inside)
                            v = (far && !lit && x < floor) ? (x - rle_lo < rle_hi - rle_lo ? (uint32_t) rle_byte : (uint32_t) o.out[x])
                                                           : (uint32_t) o.ring[at];
                        uint32_t pend = __ballot_sync(CRYO_FULL, inside);

                        if (pend)
                        {
                            uint32_t j = (lane - s_off) & 31u;      /* the lane that holds the source byte */
                            bool     have_v = !inside;

                            while (pend)
                            {
                                const uint32_t sv = __shfl_sync(CRYO_FULL, v, (int) j);
                                const uint32_t sj = __shfl_sync(CRYO_FULL, j, (int) j);

                                if (!have_v)
                                {
                                    if (!((pend >> j) & 1u))
                                    {
                                        v = sv;
                                        have_v = true;
                                    }
                                    else
                                        j = sj;
                                }
                                pend = __ballot_sync(CRYO_FULL, !have_v);
                            }
                        }
                        if (live)
                            o.ring[p & WX_RMASK] = (uint8_t) v;
                        __syncwarp();
                    }
                    o.pos = end;
                    wx_drain(o, lane);
                    k0 = k1;
                }
                lpos = base_lit + __shfl_sync(CRYO_FULL, lcum, 31);
            }
            L.pos = lpos;
            if (err != ST_OK)
                break;
        }
        /* literals left after the last sequence */
        const uint32_t rest = L.n - L.pos;

        if (rest > cap - o.pos)
        {
            err = ST_OUTPUT;
            break;
        }
        if (o.pos + rest - block_start > ZS_MAXBLOCK)
        {
            err = ST_FORMAT;
            break;
        }
        zsw_lits_emit(o, L, rest, lane);
    }
    wx_drain_all(o, lane);
    if (err == ST_OK && a.fr[(size_t) f * ZP_FF + 2] && o.pos != a.fr[(size_t) f * ZP_FF + 1])
        err = ST_SIZE;
    if (lane == 0)
    {
        if (err == ST_OK)
        {
            a.out_size[f] = o.pos;
            a.status[f] = ST_OK;
            if (a.pf_expect)
                a.pf_expect[f] = skipped;       /* what stage 0 has to have published when both kernels are done */
        }
        else
            a.flag[f] = 1;                      /* the warp-per-frame decoder rules on it */
    }
}

/* after stage 0 and stage 4 (one thread per job slot, n x ZP_JOBS): a frame whose blocks / runs stage 0 did not all publish (it gave up waiting for a job that came
 * late) goes to the warp-per-frame decoder like any other frame the pipeline declined */
CRYO_DEV void zp_stage5_check(const ZpArgs &a, uint32_t t)
{
    if (!a.pf_expect)
        return;
    /* thread t: frame t's blocks, and job slot t */
    if (t < a.n && a.flag[t] == 0 && a.pf_expect[t] != 0 && a.pf_done[t] < a.pf_expect[t])
        a.flag[t] = 1;
    if (t < a.n * ZP_JOBS && t < reinterpret_cast<const uint32_t *>(a.seq_alloc)[ZPC_JOB_TAIL])
    {
        const uint32_t *job = a.jobs + 4u * (size_t) t;

        if (!(job[3] & ZP_JOB_DONE))
            a.flag[job[0]] = 1;
    }
}
