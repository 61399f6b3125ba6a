/*
 * zstd_decode_c.cuh -- stage 4 of the zstd pipeline (zstd_decode_p.cuh) with ONE CTA per frame:
 * the latency path, and the path for frames with many sequences (a dense low-cardinality cryo
 * block has 103 K of them; one warp walking them in order, k_zp_execute, is where such a batch
 * spent 26 of its 31 ms).
 *
 * Stages 1-3 have left, per zstd block, the decoded literals and an array of (literal length,
 * match length, offset value) records.  Nothing in those records depends on earlier output, so
 * the whole CTA takes them 1 024 at a time through the cooperative executor (cryo_cx.cuh).  What
 * is serial in the format is the repeat-offset history (RFC 8878 3.1.1.5: an offset value of
 * 1..3 names one of the last three offsets).  Each sequence is a small function on that history
 * -- push a new offset, or permute the three -- and such functions compose, so the history after
 * every sequence is a prefix "sum" over the chunk with composition as the operator
 * (zc_compose): log2(1024) steps instead of 1 024.
 *
 * Any format violation raises the frame's flag; the one-warp-per-frame decoder then decodes the
 * frame from scratch and names the status, exactly as for the other stages of the pipeline.
 */
#pragma once
#include "cryo_cx.cuh"
#include "zstd_decode_p.cuh"

/* shared memory of one CTA */
#define ZC_OFF_RING   0u
#define ZC_OFF_PAT    (ZC_OFF_RING + CX_RING)
#define ZC_OFF_DL     (ZC_OFF_PAT + CX_PAT + 32u)
#define ZC_OFF_CXSH   (ZC_OFF_DL + 2u * CX_SPAN)
#define ZC_OFF_REP    ((ZC_OFF_CXSH + (uint32_t) sizeof(CxSh) + 15u) & ~15u)
#define ZC_SMEM       (ZC_OFF_REP + (3u * (CX_WARPS + 1u) + 8u) * 4u)

/*
 * A function on the repeat-offset history (r0, r1, r2): three slots, each either a constant or
 * "input slot s minus d".  Packed: bits 30-31 = s (3: constant), bits 0-29 = value or d.
 */
struct ZcRep
{
    uint32_t s[3];
};

#define ZC_IN(k)  ((uint32_t) (k) << 30)
#define ZC_CONST  (3u << 30)
#define ZC_VAL(x) ((x) & 0x3FFFFFFFu)

CRYO_DEV ZcRep zc_identity()
{
    ZcRep t;

    t.s[0] = ZC_IN(0);
    t.s[1] = ZC_IN(1);
    t.s[2] = ZC_IN(2);
    return t;
}

/* the function of one sequence: offset value ov (>= 1), ll0 = its literal length is zero */
CRYO_DEV ZcRep zc_of_sequence(uint32_t ov, bool ll0)
{
    ZcRep t;

    if (ov > 3u)
    {
        t.s[0] = ZC_CONST | ZC_VAL(ov - 3u);
        t.s[1] = ZC_IN(0);
        t.s[2] = ZC_IN(1);
        return t;
    }
    const uint32_t idx = ov - 1u + (ll0 ? 1u : 0u);

    if (idx == 0)
        return zc_identity();
    t.s[0] = idx == 1 ? ZC_IN(1) : idx == 2 ? ZC_IN(2) : (ZC_IN(0) | 1u);      /* idx 3: r0 - 1 */
    t.s[1] = ZC_IN(0);
    t.s[2] = idx == 1 ? ZC_IN(2) : ZC_IN(1);
    return t;
}

/* first a, then b */
CRYO_DEV ZcRep zc_compose(const ZcRep &a, const ZcRep &b)
{
    ZcRep c;

#pragma unroll
    for (int k = 0; k < 3; k++)
    {
        const uint32_t bs = b.s[k] >> 30;

        if (bs == 3u)
            c.s[k] = b.s[k];
        else
        {
            const uint32_t as = a.s[bs];

            c.s[k] = (as >> 30) == 3u ? (ZC_CONST | ZC_VAL(ZC_VAL(as) - ZC_VAL(b.s[k])))
                                      : ((as & ZC_CONST) | ZC_VAL(ZC_VAL(as) + ZC_VAL(b.s[k])));
        }
    }
    return c;
}

CRYO_DEV uint32_t zc_apply(uint32_t slot, uint32_t r0, uint32_t r1, uint32_t r2)
{
    const uint32_t s = slot >> 30;

    return s == 3u ? ZC_VAL(slot) : (s == 0 ? r0 : s == 1 ? r1 : r2) - ZC_VAL(slot);
}

/* inclusive scan of the sequence functions over the CTA; ws: 3 * (CX_WARPS + 1) words of shared memory */
CRYO_DEV ZcRep zc_scan(ZcRep t, uint32_t *ws, uint32_t tid)
{
    const uint32_t lane = tid & 31u, warp = tid >> 5;

#pragma unroll
    for (uint32_t d = 1; d < 32; d <<= 1)
    {
        ZcRep p;

        p.s[0] = __shfl_up_sync(CRYO_FULL, t.s[0], d);
        p.s[1] = __shfl_up_sync(CRYO_FULL, t.s[1], d);
        p.s[2] = __shfl_up_sync(CRYO_FULL, t.s[2], d);
        if (lane >= d)
            t = zc_compose(p, t);
    }
    if (lane == 31)
    {
        ws[3 * warp] = t.s[0];
        ws[3 * warp + 1] = t.s[1];
        ws[3 * warp + 2] = t.s[2];
    }
    __syncthreads();
    if (warp == 0)
    {
        ZcRep w = zc_identity();

        if (lane < CX_WARPS)
        {
            w.s[0] = ws[3 * lane];
            w.s[1] = ws[3 * lane + 1];
            w.s[2] = ws[3 * lane + 2];
        }
#pragma unroll
        for (uint32_t d = 1; d < 32; d <<= 1)
        {
            ZcRep p;

            p.s[0] = __shfl_up_sync(CRYO_FULL, w.s[0], d);
            p.s[1] = __shfl_up_sync(CRYO_FULL, w.s[1], d);
            p.s[2] = __shfl_up_sync(CRYO_FULL, w.s[2], d);
            if (lane >= d)
                w = zc_compose(p, w);
        }
        /* exclusive: warp w needs the composition of the warps before it */
        ZcRep e;

        e.s[0] = __shfl_up_sync(CRYO_FULL, w.s[0], 1);
        e.s[1] = __shfl_up_sync(CRYO_FULL, w.s[1], 1);
        e.s[2] = __shfl_up_sync(CRYO_FULL, w.s[2], 1);
        if (lane == 0)
            e = zc_identity();
        __syncwarp();
        if (lane < CX_WARPS)
        {
            ws[3 * lane] = e.s[0];
            ws[3 * lane + 1] = e.s[1];
            ws[3 * lane + 2] = e.s[2];
        }
    }
    __syncthreads();
    ZcRep pre;

    pre.s[0] = ws[3 * warp];
    pre.s[1] = ws[3 * warp + 1];
    pre.s[2] = ws[3 * warp + 2];
    return zc_compose(pre, t);
}

/* stage 4 body: one CTA, frame f.  smem: ZC_SMEM bytes. */
CRYO_DEV void zp_stage4_cx(const ZpArgs &a, uint32_t f, uint8_t *smem, uint32_t tid)
{
    const uint32_t nb = a.fr[(size_t) f * ZP_FF], cap = a.cap;
    const uint8_t *in = a.src + a.src_off[f];
    Cx       cx;
    CxSh    *sh = reinterpret_cast<CxSh *>(smem + ZC_OFF_CXSH);
    uint32_t *rws = reinterpret_cast<uint32_t *>(smem + ZC_OFF_REP);
    uint32_t *carry = rws + 3u * (CX_WARPS + 1u);       /* rep0..2 after the sequences executed so far, literal bytes used */
    int      err = ST_OK;
    uint32_t rep0 = 1, rep1 = 4, rep2 = 8;

    cx_init(cx, a.dst + (size_t) f * a.dst_stride, cap, smem + ZC_OFF_RING, smem + ZC_OFF_PAT,
            reinterpret_cast<uint16_t *>(smem + ZC_OFF_DL), sh);
#if defined(CX_PROF) && !defined(CRYO_EMU)
    if (threadIdx.x == 0 && blockIdx.x == 0)
        cx_prof_last = clock64();
#endif
    for (uint32_t j = 0; j < nb && err == ST_OK; j++)
    {
        const uint32_t *b = a.blk + ((size_t) f * ZP_MAXB + j) * ZP_BF;
        const uint32_t off = b[ZPB_OFF], bsize = b[ZPB_BSIZE], kind = b[ZPB_KIND], type = kind & 3u;

        if (type < 2)
        {
            /* Raw / RLE block: one literal run */
            if (bsize == 0)
                continue;
            __syncthreads();
            uint32_t cum = tid == 0 ? bsize : 0u, dummy = 0;

            cx_scan2(sh, cum, dummy, tid);
            if (cx_chunk(cx, 1, tid == 0 ? bsize : 0u, 0, 0, in + off, type == 1 ? (int) in[off] : -1, cum, ST_OK, tid) == 0)
                err = cx.err;
            continue;
        }
        const uint32_t lt = (kind >> 2) & 3u, regen = b[ZPB_REGEN], nseq = b[ZPB_NSEQ];
        const uint32_t block_start = cx.pos;
        const uint8_t *lit_base = lt >= 2 ? a.lit + (size_t) f * a.lit_stride + b[ZPB_LITPOS] : in + off + b[ZPB_LHDR];
        const int      rle = lt == 1 ? (int) lit_base[0] : -1;
        const uint64_t *sq = a.seq + a.seqbase[f] + b[ZPB_SEQPOS];
        uint32_t lpos = 0, c0 = 0;

        while (c0 < nseq && err == ST_OK)
        {
            const uint32_t n = nseq - c0 < CX_THREADS ? nseq - c0 : CX_THREADS;
            const bool     valid = tid < n;
            const uint64_t r = valid ? sq[c0 + tid] : 0ull;
            const uint32_t ll = (uint32_t) r & 0x1FFFFu, ml = (uint32_t) (r >> 17) & 0x3FFFFu;
            const uint32_t ov = (uint32_t) (r >> 35);
            uint32_t cum = valid ? ll + ml : 0u, lcum = valid ? ll : 0u;

            __syncthreads();
            CXP(5)
            cx_scan2(sh, cum, lcum, tid);
            CXP(6)
            /* repeat offsets: the history after every sequence of the chunk */
            const ZcRep t = zc_scan(valid ? zc_of_sequence(ov, ll == 0) : zc_identity(), rws, tid);
            const uint32_t n0 = zc_apply(t.s[0], rep0, rep1, rep2);     /* = this sequence's offset */
            const uint32_t n1 = zc_apply(t.s[1], rep0, rep1, rep2), n2 = zc_apply(t.s[2], rep0, rep1, rep2);
            int            pre = ST_OK;

            if (valid)
            {
                const uint32_t epos = cx.pos + cum;

                /* literals beyond the block's, a block regenerating more than 128 KiB, offset 0 (a repeat
                 * offset of 1 minus 1), a sequence without a match: the fallback decoder rules on all of them */
                if (lpos + lcum > regen || epos - block_start > ZS_MAXBLOCK || n0 == 0 || ml == 0 || ov == 0)
                    pre = ST_FORMAT;
            }
            const uint8_t *lp = rle >= 0 ? lit_base : lit_base + (lpos + lcum - ll);
            CXP(7)
            const uint32_t k = cx_chunk(cx, n, ll, ml, n0, lp, rle, cum, pre, tid);

            if (k == 0)
            {
                err = cx.err;
                break;
            }
            if (tid == k - 1u)
            {
                carry[0] = n0;
                carry[1] = n1;
                carry[2] = n2;
                carry[3] = lcum;
            }
            __syncthreads();
            rep0 = carry[0];
            rep1 = carry[1];
            rep2 = carry[2];
            lpos += carry[3];
            c0 += k;
        }
        if (err != ST_OK)
            break;
        /* literals left after the last sequence */
        const uint32_t rest = regen - lpos;

        if (rest > cap - cx.pos || cx.pos + rest - block_start > ZS_MAXBLOCK)
        {
            err = ST_FORMAT;
            break;
        }
        if (rest)
        {
            __syncthreads();
            uint32_t cum = tid == 0 ? rest : 0u, dummy = 0;

            cx_scan2(sh, cum, dummy, tid);
            if (cx_chunk(cx, 1, tid == 0 ? rest : 0u, 0, 0, rle >= 0 ? lit_base : lit_base + lpos, rle, cum, ST_OK, tid) == 0)
                err = cx.err;
        }
    }
    cx_finish(cx, tid);
    CXP(14)
    if (err == ST_OK && a.fr[(size_t) f * ZP_FF + 2] && cx.pos != a.fr[(size_t) f * ZP_FF + 1])
        err = ST_SIZE;
    if (tid == 0)
    {
        if (err == ST_OK)
        {
            a.out_size[f] = cx.pos;
            a.status[f] = ST_OK;
        }
        else
            a.flag[f] = 1;              /* the warp-per-frame decoder rules on it */
    }
}
