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
#define ZP_BF     18u           /* u32 fields per block descriptor */
#define ZP_PREFILL_MIN 2048u     /* raw / RLE blocks at least this long are written ahead by stage 0 */
#define ZP_FF     4u            /* u32 fields per frame descriptor: nblk, fcs, has_fcs, - */

enum
{
    ZPB_OFF = 0,                /* frame-relative offset of the block content */
    ZPB_BSIZE,                  /* content bytes (raw, compressed) or run length (RLE) */
    ZPB_KIND,                   /* bits 0-1 block type, 2-3 literal type, 4 four streams, 8-15 modes */
    ZPB_LHDR,                   /* literals section header bytes */
    ZPB_REGEN,                  /* literals regenerated size */
    ZPB_LCSIZE,                 /* literals compressed size (tree description included) */
    ZPB_HDOFF,                  /* frame-relative offset of the Huffman tree description in force */
    ZPB_HDLEFT,                 /* bytes readable there */
    ZPB_NSEQ,
    ZPB_SEQOFF,                 /* block-relative offset of the first table description */
    ZPB_LITPOS,                 /* byte offset in the frame's literal area (multiple of 16) */
    ZPB_SEQPOS,                 /* entry offset in the frame's sequence area */
    ZPB_SPECPOS,                /* raw / RLE: output position stage 0 assumed and wrote at, or ~0u */
    ZPB_HDBLK,                  /* index of the block whose Huffman tree is in force */
    ZPB_HINFO,                  /* written by stage 2a: table log | description bytes << 8 (0: failed) */
    ZPB_SLOGS,                  /* written by stage 3a: ll_log | of_log << 8 | ml_log << 16 | 1 << 31 */
    ZPB_BITOFF                  /* written by stage 3a: block-relative offset of the sequence bitstream */
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
    uint64_t spec = 0;                          /* output position if every earlier Compressed block is full */
    const uint64_t spec_lim = fcs_bytes ? fcs : 0;
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
        b[ZPB_SPECPOS] = (type < 2 && bsize >= ZP_PREFILL_MIN && spec + bsize <= spec_lim) ? (uint32_t) spec : ~0u;
        spec += type < 2 ? bsize : ZS_MAXBLOCK;
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
            b[ZPB_NSEQ] = nseq;
            b[ZPB_SEQOFF] = sp;
            b[ZPB_LITPOS] = lit_total;
            b[ZPB_SEQPOS] = seq_total;
            if (lt >= 2)
                lit_total += (regen + 15u) & ~15u;
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
    *seq_total_out = seq_total;
    return true;
}

/* stage 1 body: the calling lane owns frame f */
CRYO_DEV void zp_stage1(const ZpArgs &a, uint32_t f)
{
    uint32_t *fr = a.fr + (size_t) f * ZP_FF;

    fr[0] = 0;
    a.flag[f] = 0;
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
    }
    if (!ok)
    {
        fr[0] = 0;
        a.flag[f] = 1;
    }
}

/* ------------------------------------------------- stage 0: raw / RLE blocks ahead ---- */

/*
 * libzstd cuts a frame into full 128 KiB blocks (only the last one is short), so the output
 * position of a Raw or RLE block is known from the headers alone if that holds.  Stage 0
 * writes those blocks at the assumed position while stages 2 and 3 run (they are bound by
 * latency, this is bound by HBM); stage 4 skips a block when it arrives at exactly that
 * position and writes it itself otherwise (anything stage 0 wrote is then overwritten).
 * One CTA per (frame, block index).
 */
CRYO_DEV void zp_stage0(const ZpArgs &a, uint32_t f, uint32_t j, uint32_t tid, uint32_t nthr)
{
    if (f >= a.n || a.fr[(size_t) f * ZP_FF] <= j)
        return;
    const uint32_t *b = a.blk + ((size_t) f * ZP_MAXB + j) * ZP_BF;
    const uint32_t spec = b[ZPB_SPECPOS];

    if ((b[ZPB_KIND] & 3u) >= 2u || spec == ~0u)
        return;
    const uint8_t *in = a.src + a.src_off[f] + b[ZPB_OFF];
    uint8_t *dst = a.dst + (size_t) f * a.dst_stride + spec;

    if ((b[ZPB_KIND] & 3u) == 0)
        team_copy(dst, in, b[ZPB_BSIZE], tid, nthr);
    else
        team_fill_byte(dst, in[0], b[ZPB_BSIZE], tid, nthr);
}

/* --------------------------------------------------------------- stage 2: literals ---- */

/*
 * 2a: one warp per Huffman tree description: weights (lane 0), table filled by the warp
 *     straight into the block's slot of huftab (global, u16[2048] per slot); log and
 *     description length go into the block descriptor.
 * 2b: one warp per (8 frames x block index), 32 lanes = 8 blocks x 4 streams, tables read
 *     through L1.  No shared memory: the stage co-resides with anything.
 */
#define ZP2A_WARPS      8u
#define ZP2A_PER_WARP   (704u + 512u + 128u)    /* weights work | symstart u16[256] | rankc u32[32] */
#define ZP2A_SMEM       (ZP2A_WARPS * ZP2A_PER_WARP)

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

/* stage 2a body: one warp, block j of frame f */
CRYO_DEV void zp_stage2a(const ZpArgs &a, uint32_t f, uint32_t j, uint8_t *smem, uint32_t lane)
{
    if (f >= a.n || a.fr[(size_t) f * ZP_FF] <= j)
        return;
    uint32_t *b = a.blk + ((size_t) f * ZP_MAXB + j) * ZP_BF;
    const uint32_t kind = b[ZPB_KIND];

    if ((kind & 3u) != 2u || ((kind >> 2) & 3u) != 2u)
        return;                                 /* treeless blocks use the slot of the block that defined the tree */
    uint32_t used = 0, nw = 0;
    int32_t  log = 0;

    if (lane == 0)
        used = zp_huf_weights(a.src + a.src_off[f] + b[ZPB_HDOFF], b[ZPB_HDLEFT], smem,
                              reinterpret_cast<uint32_t *>(smem + 256), reinterpret_cast<int16_t *>(smem + 512),
                              reinterpret_cast<uint16_t *>(smem + 544), &nw);
    used = __shfl_sync(CRYO_FULL, used, 0);
    nw = __shfl_sync(CRYO_FULL, nw, 0);
    __syncwarp();
    const bool ok = used != 0 &&
                    zsw_huf_table(smem, nw, a.huftab + ((size_t) f * ZP_MAXB + j) * 2048u,
                                  reinterpret_cast<uint16_t *>(smem + 704), reinterpret_cast<uint32_t *>(smem + 1216),
                                  &log, lane);

    if (lane == 0)
    {
        if (ok)
            b[ZPB_HINFO] = (uint32_t) log | (used << 8);
        else
            a.flag[f] = 1;
    }
}

/* stage 2b body: one warp, block index j of frames [g * ZP_G, g * ZP_G + ZP_G) */
#define ZP2B_SMEM       (ZP_G * 4096u)

CRYO_DEV void zp_stage2b(const ZpArgs &a, uint32_t g, uint32_t j, uint8_t *smem, uint32_t lane)
{
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
    /* the group's tables -> shared memory: the 4 lanes of a table copy it, 16 bytes at a time */
    if (valid && (hinfo & 0xFFu) != 0)
    {
        const uint8_t *gt = reinterpret_cast<const uint8_t *>(a.huftab + ((size_t) tf * ZP_MAXB + hb) * 2048u);
        uint8_t *st = smem + ti * 4096u;
        const uint32_t bytes = 2u << (hinfo & 0xFFu);

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
    __syncwarp();
    if (!valid)
        return;
    const int32_t  tlog = (int32_t) (hinfo & 0xFFu);
    const uint32_t own = lt == 2u ? hinfo >> 8 : 0u;
    const uint8_t *p = a.src + a.src_off[tf] + b[ZPB_OFF] + b[ZPB_LHDR] + own;
    const uint32_t left = b[ZPB_LCSIZE] - own, regen = b[ZPB_REGEN];
    uint8_t *dst = a.lit + (size_t) tf * a.lit_stride + b[ZPB_LITPOS];
    const uint16_t *huf = reinterpret_cast<const uint16_t *>(smem + ti * 4096u);
    bool ok = true;

    if (tlog == 0)
        ok = false;                             /* the tree's own stage-2a warp flagged the frame */
    else if (!((kind >> 4) & 1u))
    {
        if (s == 0)
            ok = zsw_huf_stream(huf, tlog, p, left, dst, regen);
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
            const uint32_t sn = s == 0 ? s1 : s == 1 ? s2 : s == 2 ? s3 : s4;
            const uint32_t cnt = s < 3 ? seg : regen - 3 * seg;

            ok = zsw_huf_stream(huf, tlog, p + 6 + so, sn, dst + s * seg, cnt);
        }
    }
    if (!ok)
        a.flag[tf] = 1;
}

/* -------------------------------------------------------------- stage 3: sequences ---- */

/*
 * 3a: one warp per block: the three table descriptions (lane 0 reads the counts), cells built
 *     by the warp in shared memory and copied to the block's slot of fsetab (global,
 *     LL u32[512] | OF u32[256] | ML u32[512]); logs and the bitstream offset go into the
 *     descriptor.
 * 3b: one LANE per block, 32 blocks (same block index of 32 frames) per warp, cells read
 *     through L1, (ll, ml, offset value) written to the frame's sequence area.
 */
#define ZP3_CELLS       1280u                   /* u32 cells per block slot */
#define ZP3A_WARPS      8u
#define ZP3A_PER_WARP   (ZP3_CELLS * 4u + 128u + 128u + 144u)   /* cells | counts i16[64] | next u16[64] | cum u16[66] */
#define ZP3A_SMEM       (ZP3A_WARPS * ZP3A_PER_WARP)

/* stage 3a body: one warp, block j of frame f */
CRYO_DEV void zp_stage3a(const ZpArgs &a, uint32_t f, uint32_t j, uint8_t *smem, uint32_t lane)
{
    if (f >= a.n || a.fr[(size_t) f * ZP_FF] <= j)
        return;
    uint32_t *b = a.blk + ((size_t) f * ZP_MAXB + j) * ZP_BF;
    const uint32_t kind = b[ZPB_KIND];

    if ((kind & 3u) != 2u || b[ZPB_NSEQ] == 0)
        return;
    const uint8_t *p = a.src + a.src_off[f] + b[ZPB_OFF] + b[ZPB_SEQOFF];
    uint32_t left = b[ZPB_BSIZE] - b[ZPB_SEQOFF], logs = 0;
    uint32_t *cells = reinterpret_cast<uint32_t *>(smem);
    int16_t  *counts = reinterpret_cast<int16_t *>(smem + ZP3_CELLS * 4u);
    uint16_t *next = reinterpret_cast<uint16_t *>(smem + ZP3_CELLS * 4u + 128u);
    uint16_t *cum = reinterpret_cast<uint16_t *>(smem + ZP3_CELLS * 4u + 256u);
    bool      bad = false;

#pragma unroll
    for (int t = 0; t < 3; t++)
    {
        const uint32_t mode = (kind >> (14 - 2 * t)) & 3u;
        const int max_log = t == 1 ? 8 : 9, max_sym = t == 0 ? 35 : t == 1 ? 31 : 52;
        uint32_t *cell = cells + (t == 0 ? 0u : t == 1 ? 512u : 768u);
        int32_t   logv = 0;

        if (mode == 0)
        {
            const uint32_t n = t == 1 ? 32u : 64u, o = t == 0 ? 0u : t == 1 ? 64u : 96u;

            for (uint32_t k = lane; k < n; k += 32)
                cell[k] = a.predef[o + k];
            logv = t == 1 ? 5 : 6;
        }
        else if (mode == 1)
        {
            if (left < 1 || p[0] > max_sym)
            {
                bad = true;
                break;
            }
            if (lane == 0)
                cell[0] = p[0];
            p += 1;
            left -= 1;
        }
        else
        {
            int32_t  nsym = 0, log = 0;
            uint32_t used = 0;

            if (lane == 0)
                used = fse_read_counts(p, left, max_log, max_sym, counts, &nsym, &log);
            used = __shfl_sync(CRYO_FULL, used, 0);
            nsym = __shfl_sync(CRYO_FULL, nsym, 0);
            log = __shfl_sync(CRYO_FULL, log, 0);
            if (used == 0)
            {
                bad = true;
                break;
            }
            __syncwarp();
            fse_build_table_warp(cell, counts, nsym, log, next, cum, lane);
            logv = log;
            p += used;
            left -= used;
        }
        __syncwarp();
        /* extra-bit count of every cell's code into bits 26..30 (as zsw_seq_table) */
        for (uint32_t k = lane; k < (1u << logv); k += 32)
        {
            const uint32_t c = cell[k], sym = c & 0xFFu;
            const uint32_t xb = t == 1 ? sym : (t == 0 ? CRYO_GLD(ZS_LL_PACK[sym]) : CRYO_GLD(ZS_ML_PACK[sym])) >> 24;

            cell[k] = (c & 0x03FFFFFFu) | (xb << 26);
        }
        __syncwarp();
        logs |= (uint32_t) logv << (8 * t);
    }
    if (bad)
    {
        if (lane == 0)
            a.flag[f] = 1;
        return;
    }
    /* cells -> global slot (only the live part of each table) */
    uint32_t *slot = a.fsetab + ((size_t) f * ZP_MAXB + j) * ZP3_CELLS;

    for (uint32_t k = lane; k < (1u << (logs & 0xFFu)); k += 32)
        slot[k] = cells[k];
    for (uint32_t k = lane; k < (1u << ((logs >> 8) & 0xFFu)); k += 32)
        slot[512u + k] = cells[512u + k];
    for (uint32_t k = lane; k < (1u << ((logs >> 16) & 0xFFu)); k += 32)
        slot[768u + k] = cells[768u + k];
    if (lane == 0)
    {
        b[ZPB_SLOGS] = logs | 0x80000000u;
        b[ZPB_BITOFF] = b[ZPB_BSIZE] - left;
    }
    __syncwarp();
}

/* one lane walks one block's sequence bitstream (RFC 8878 3.1.1.3.2.1.2) */
CRYO_DEV bool zp_seq_walk(const uint32_t *llt, const uint32_t *oft, const uint32_t *mlt, uint32_t ll_log,
                          uint32_t of_log, uint32_t ml_log, const uint8_t *p, uint32_t n, uint32_t nseq,
                          uint64_t *out)
{
    BitsBack bb;

    if (!bb_init(bb, p, n))
        return false;
    bb_refill(bb);
    uint32_t sl = bb_read(bb, ll_log);
    uint32_t so = bb_read(bb, of_log);
    uint32_t sm = bb_read(bb, ml_log);

    if (bb.remaining < 0)
        return false;
#ifndef CRYO_EMU
    /* L1 fills by 32-byte sector and a miss stalls every lane of the warp: ask for the sector
     * three below the read position once per sequence */
    if (bb.cur >= bb.start + 32)
        asm volatile("prefetch.global.L1 [%0];" ::"l"(bb.cur - 32));
    if (bb.cur >= bb.start + 64)
        asm volatile("prefetch.global.L1 [%0];" ::"l"(bb.cur - 64));
#endif
    for (uint32_t i = 0; i < nseq; i++)
    {
#ifndef CRYO_EMU
        if (bb.cur >= bb.start + 96)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(bb.cur - 96));
#endif
        const uint32_t cl = llt[sl], co = oft[so], cm = mlt[sm];
        const uint32_t xo = co >> 26, xm = cm >> 26, xl = cl >> 26;

        if (xo > 27)
            return false;                       /* beyond any window ZSTD_decompress accepts */
        bb_refill(bb);
        const uint32_t ov = (1u << xo) + bb_read(bb, xo);

        bb_refill(bb);
        const uint32_t ml = (CRYO_GLD(ZS_ML_PACK[cm & 0xFFu]) & 0xFFFFFFu) + bb_read(bb, xm);
        const uint32_t ll = (CRYO_GLD(ZS_LL_PACK[cl & 0xFFu]) & 0xFFFFFFu) + bb_read(bb, xl);

        if (i + 1 < nseq)
        {
            bb_refill(bb);
            sl = ((cl >> 16) & 0x3FFu) + bb_read(bb, (cl >> 8) & 0xFFu);
            sm = ((cm >> 16) & 0x3FFu) + bb_read(bb, (cm >> 8) & 0xFFu);
            so = ((co >> 16) & 0x3FFu) + bb_read(bb, (co >> 8) & 0xFFu);
        }
        if (bb.remaining < 0)
            return false;
        out[i] = (uint64_t) ll | ((uint64_t) ml << 17) | ((uint64_t) ov << 35);
    }
    return bb.remaining == 0;
}

/*
 * stage 3b body: one warp, block index j of frames [g * ZP_G, g * ZP_G + ZP_G); lane i < ZP_G
 * walks frame g * ZP_G + i.  The live cells of the three tables are packed LL | OF | ML into
 * CELLS u32 of shared memory per block.  The kernel is launched once per size class (small:
 * what libzstd emits for sparse blocks, 6/6/7-bit tables; large: the 9/8/9-bit maximum) and a
 * warp runs in the smallest class that holds all its blocks, so the small class keeps many
 * groups resident per SM.
 */
#define ZP3B_SMALL      320u
#define ZP3B_LARGE      ZP3_CELLS

template <uint32_t CELLS, uint32_t BELOW>      /* this launch takes groups needing > BELOW and <= CELLS cells */
CRYO_DEV void zp_stage3b(const ZpArgs &a, uint32_t g, uint32_t j, uint8_t *smem, uint32_t lane)
{
    const uint32_t f = g * ZP_G + lane;
    const uint32_t *b = nullptr;
    uint32_t logs = 0, nseq = 0, need = 0;
    bool     valid = false, failed = false;

    if (lane < ZP_G && f < a.n && a.fr[(size_t) f * ZP_FF] > j && a.flag[f] == 0)
    {
        b = a.blk + ((size_t) f * ZP_MAXB + j) * ZP_BF;
        nseq = b[ZPB_NSEQ];
        if ((b[ZPB_KIND] & 3u) == 2u && nseq != 0)
        {
            logs = b[ZPB_SLOGS];
            valid = (logs & 0x80000000u) != 0;
            failed = !valid;                    /* the block's stage-3a warp flagged the frame */
            need = valid ? (1u << (logs & 0xFFu)) + (1u << ((logs >> 8) & 0xFFu)) + (1u << ((logs >> 16) & 0x7Fu)) : 0u;
        }
    }
    const uint32_t gneed = __reduce_max_sync(CRYO_FULL, need);

    if (gneed == 0 || gneed > CELLS || gneed <= BELOW)
        return;
    (void) failed;
    /* tables -> shared memory, the whole warp per block */
    for (uint32_t m = __ballot_sync(CRYO_FULL, valid); m; m &= m - 1)
    {
        const int      i = __ffs((int) m) - 1;
        const uint32_t li = __shfl_sync(CRYO_FULL, logs, i);
        const uint32_t nl = 1u << (li & 0xFFu), no = 1u << ((li >> 8) & 0xFFu), nm = 1u << ((li >> 16) & 0x7Fu);
        const uint32_t *slot = a.fsetab + ((size_t) (g * ZP_G + (uint32_t) i) * ZP_MAXB + j) * ZP3_CELLS;
        uint32_t *cells = reinterpret_cast<uint32_t *>(smem) + (uint32_t) i * CELLS;

        for (uint32_t k = lane; k < nl; k += 32)
            cells[k] = slot[k];
        for (uint32_t k = lane; k < no; k += 32)
            cells[nl + k] = slot[512u + k];
        for (uint32_t k = lane; k < nm; k += 32)
            cells[nl + no + k] = slot[768u + k];
    }
    __syncwarp();
    if (!valid)
        return;
    const uint32_t ll_log = logs & 0xFFu, of_log = (logs >> 8) & 0xFFu, ml_log = (logs >> 16) & 0x7Fu;
    const uint32_t *cells = reinterpret_cast<const uint32_t *>(smem) + lane * CELLS;
    const uint32_t bitoff = b[ZPB_BITOFF];

    if (!zp_seq_walk(cells, cells + (1u << ll_log), cells + (1u << ll_log) + (1u << of_log), ll_log, of_log, ml_log,
                     a.src + a.src_off[f] + b[ZPB_OFF] + bitoff, b[ZPB_BSIZE] - bitoff, nseq,
                     a.seq + a.seqbase[f] + b[ZPB_SEQPOS]))
        a.flag[f] = 1;
}

/* ---------------------------------------------------------------- stage 4: execute ---- */

#define ZP4_WARPS       8u
#define ZP4_THREADS     (32u * ZP4_WARPS)
#define ZP4_PER_WARP    (WX_RING + ZSW_LITWIN)
#define ZP4_SMEM        (ZP4_WARPS * ZP4_PER_WARP)

/* stage 4 body: one warp, frame f */
CRYO_DEV void zp_stage4(const ZpArgs &a, uint32_t f, uint8_t *smem, uint32_t lane)
{
    if (f >= a.n || a.methods[f] != ZP_METHOD_ZSTD || a.flag[f] != 0)
        return;
    const uint32_t nb = a.fr[(size_t) f * ZP_FF], cap = a.cap;
    const uint8_t *in = a.src + a.src_off[f];
    WOut     o;
    int      err = ST_OK;
    uint32_t rep0 = 1, rep1 = 4, rep2 = 8;

    wx_init(o, a.dst + (size_t) f * a.dst_stride, cap, smem);
    for (uint32_t j = 0; j < nb && err == ST_OK; j++)
    {
        const uint32_t *b = a.blk + ((size_t) f * ZP_MAXB + j) * ZP_BF;
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
            if (b[ZPB_SPECPOS] == o.pos)
            {
                /* stage 0 wrote this block here already */
                wx_drain_all(o, lane);
                wx_after_bulk(o, bsize, lane);
                continue;
            }
            if (type == 0)
                wx_literals(o, in + off, bsize, lane);
            else
                wx_fill_byte(o, in[off], bsize, lane);
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

            for (uint32_t done = 0; done < nseq && err == ST_OK; done += 32)
            {
                const uint32_t g = nseq - done < 32u ? nseq - done : 32u;
                const uint64_t cur = nxt;

                if (done + 32u + lane < nseq)
                    nxt = sq[done + 32u + lane];
                const uint32_t my_ll = (uint32_t) cur & 0x1FFFFu, my_ml = (uint32_t) (cur >> 17) & 0x3FFFFu;
                const uint32_t my_ov = (uint32_t) (cur >> 35);

                for (uint32_t k = 0; k < g; k++)
                {
                    const uint32_t ov = __shfl_sync(CRYO_FULL, my_ov, (int) k);
                    const uint32_t ml = __shfl_sync(CRYO_FULL, my_ml, (int) k);
                    const uint32_t ll = __shfl_sync(CRYO_FULL, my_ll, (int) k);
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
                        const uint32_t idx = ov - 1 + (ll == 0 ? 1u : 0u);

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
                    const uint32_t mpos = o.pos + ll, epos = mpos + ml;     /* < 2^28: no wrap */

                    if ((ll > regen - lpos) | (epos > cap) | (epos - block_start > ZS_MAXBLOCK) |
                        (moff - 1u >= mpos))
                    {
                        err = ll > regen - lpos ? ST_FORMAT
                              : epos > cap ? ST_OUTPUT
                              : epos - block_start > ZS_MAXBLOCK ? ST_FORMAT : ST_OFFSET;
                        break;
                    }
                    /* fast path: literal run and non-overlapping match of up to 64 bytes each,
                     * literals in the window, match source in the ring */
                    const uint32_t lip = L.delta + lpos;

                    if (ll <= 64u && ml <= 64u && !L.rle && moff >= ml && moff <= WX_RING - 64u &&
                        mpos - moff >= o.lo)
                    {
                        if (ll)
                        {
                            if (!L.wvalid || lip + ll > L.wbase + ZSW_LITWIN || lip < L.wbase)
                                zsw_lits_fill(L, lip, lane);
                            if (lane < ll)
                                o.ring[(o.pos + lane) & WX_RMASK] = L.win[lip - L.wbase + lane];
                            if (lane + 32u < ll)
                                o.ring[(o.pos + lane + 32u) & WX_RMASK] = L.win[lip - L.wbase + lane + 32u];
                            __syncwarp();
                        }
                        if (lane < ml)
                            o.ring[(mpos + lane) & WX_RMASK] = o.ring[(mpos - moff + lane) & WX_RMASK];
                        if (lane + 32u < ml)
                            o.ring[(mpos + lane + 32u) & WX_RMASK] = o.ring[(mpos - moff + lane + 32u) & WX_RMASK];
                        o.pos = epos;
                        lpos += ll;
                        __syncwarp();
                        if (o.pos - o.flushed >= WX_DRAIN)
                            wx_drain(o, lane);
                        continue;
                    }
                    L.pos = lpos;
                    zsw_lits_emit(o, L, ll, lane);
                    lpos += ll;
                    wx_match(o, moff, ml, lane);
                }
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
        }
        else
            a.flag[f] = 1;                      /* the warp-per-frame decoder rules on it */
    }
}
