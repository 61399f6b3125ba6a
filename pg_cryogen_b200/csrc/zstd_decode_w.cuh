/*
 * zstd_decode_w.cuh -- batched zstd frame decompression, ONE WARP per cryo block
 * (the throughput path; zstd_decode.cuh is the one-CTA-per-block variant and
 * holds the bit readers and table readers both variants share).
 *
 * Replaces ZSTD_decompress as called at reference compression.c:116.  Same
 * coverage of RFC 8878 as zstd_decode.cuh.  What changes is the mapping to the
 * machine: the FSE sequence stream of a block is one serial chain whatever is
 * thrown at it, so instead of parking three idle warps next to the one that
 * walks it, every warp walks its own frame and the per-warp shared-memory
 * footprint is kept at ~11 KB so that ~20 frames are in flight per SM:
 *
 *   entropy phase  (ring idle, its 6 KB hold the Huffman table + scratch)
 *       Huffman tree -> table (warp-parallel fill), the literal streams decoded one
 *       lane per stream into a global scratch line (stays in L2), the three FSE
 *       tables built warp-parallel (spread by closed form, state numbering by
 *       __match_any_sync ranks, 32 cells per step)
 *   sequence phase (ring + 1 KB literal window live in the same 6 KB)
 *       all lanes decode the FSE bitstream redundantly (state is warp-uniform)
 *       and execute each sequence through cryo_wexec.cuh as it appears
 */
#pragma once
#include "cryo_wexec.cuh"
#include "zstd_format.cuh"

#define ZSW_WARPS     8
#define ZSW_THREADS   (32 * ZSW_WARPS)
#define ZSW_CTAS_PER_SM 3
#define ZSW_LITWIN    512u
#define ZSW_SEQWIN    512u
/*
 * per-warp shared memory, 9 216 bytes (24 frames in flight per SM):
 *   [0, 4096)     literal phase : Huffman table u16[2048]
 *                 table phase   : FSE build scratch at ZSW_OFF_FSEWORK (the Huffman table is dead)
 *                 sequence phase: ring [0, 2048) | literal window | sequence-bitstream window
 *   [4096, 9216)  FSE cells LL u32[512] | OF u32[256] | ML u32[512]; the Huffman build scratch
 *                 overlays the LL cells, so a table reused through Repeat_Mode is rebuilt from its
 *                 remembered description
 */
#define ZSW_OFF_RING    0
#define ZSW_OFF_LITWIN  WX_RING                               /* 2048 */
#define ZSW_OFF_SEQWIN  (ZSW_OFF_LITWIN + ZSW_LITWIN)         /* 2560: 16 B zero pad + window + 16 */
#define ZSW_OFF_HUF     0
#define ZSW_OFF_FSEWORK 2048                                  /* ZW_COUNTS.. offsets land in [3200, 3900) */
#define ZSW_OFF_LL      4096
#define ZSW_OFF_OF      (ZSW_OFF_LL + 2048)
#define ZSW_OFF_ML      (ZSW_OFF_OF + 1024)
#define ZSW_OFF_HUFWORK ZSW_OFF_LL
#define ZSW_PER_WARP    (ZSW_OFF_ML + 2048)     /* 9216 */
#define ZSW_SMEM        (ZSW_WARPS * ZSW_PER_WARP)
#define ZSW_PREDEF_CELLS (64 + 32 + 64)         /* LL, OF, ML predefined tables */

#if ZSW_OFF_SEQWIN + ZSW_SEQWIN + 32 > 4096
#error "ring + windows must fit the 4 KB they share with the Huffman table"
#endif

/* code -> baseline | extra bits << 24 (RFC 8878 3.1.1.3.2.1.1).  Looked up with a different
 * index per lane, so they live in global memory (read-only cache), not in the constant bank
 * where divergent indices serialise. */
#ifdef CRYO_EMU
#define CRYO_GTABLE static const
#define CRYO_GLD(x) (x)
#else
#define CRYO_GTABLE __device__ const
#define CRYO_GLD(x) __ldg(&(x))
#endif
CRYO_GTABLE uint32_t ZS_LL_PACK[36] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15,
    16 | (1u << 24), 18 | (1u << 24), 20 | (1u << 24), 22 | (1u << 24), 24 | (2u << 24), 28 | (2u << 24),
    32 | (3u << 24), 40 | (3u << 24), 48 | (4u << 24), 64 | (6u << 24), 128 | (7u << 24), 256 | (8u << 24),
    512 | (9u << 24), 1024 | (10u << 24), 2048 | (11u << 24), 4096 | (12u << 24), 8192 | (13u << 24),
    16384 | (14u << 24), 32768 | (15u << 24), 65536 | (16u << 24)};
CRYO_GTABLE uint32_t ZS_ML_PACK[53] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20,
    21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35 | (1u << 24), 37 | (1u << 24),
    39 | (1u << 24), 41 | (1u << 24), 43 | (2u << 24), 47 | (2u << 24), 51 | (3u << 24), 59 | (3u << 24),
    67 | (4u << 24), 83 | (4u << 24), 99 | (5u << 24), 131 | (7u << 24), 259 | (8u << 24), 515 | (9u << 24),
    1027 | (10u << 24), 2051 | (11u << 24), 4099 | (12u << 24), 8195 | (13u << 24), 16387 | (14u << 24),
    32771 | (15u << 24), 65539 | (16u << 24)};

CRYO_DEV uint32_t zsw_match_any(uint32_t v)
{
#ifdef CRYO_EMU
    uint32_t m = 0;

    for (int l = 0; l < 32; l++)
        if (__shfl_sync(CRYO_FULL, v, l) == v)
            m |= 1u << l;
    return m;
#else
    return __match_any_sync(CRYO_FULL, v);
#endif
}

/*
 * Warp-parallel FSE decoding-table build (RFC 8878 4.1.1), same result as the
 * serial fse_build_table: cell = symbol | nbits << 8 | base << 16.
 * counts[nsym] in shared memory; next[64] and cum[65] are shared scratch.
 */
CRYO_DEV void fse_build_table_warp(uint32_t *cell, const int16_t *counts, int nsym, int log,
                                   uint16_t *next, uint16_t *cum, uint32_t lane)
{
    const uint32_t size = 1u << log, mask = size - 1u;
    const uint32_t step = (size >> 1) + (size >> 3) + 3u;
    const uint32_t lt = (1u << lane) - 1u;

    /* low-probability symbols sit at the top of the table, one cell each */
    int      cA = (int) lane < nsym ? counts[lane] : 0;
    int      cB = (int) lane + 32 < nsym ? counts[lane + 32] : 0;
    uint32_t mA = __ballot_sync(CRYO_FULL, cA == -1), mB = __ballot_sync(CRYO_FULL, cB == -1);
    uint32_t nlowA = (uint32_t) __popc(mA), nlow = nlowA + (uint32_t) __popc(mB);
    const uint32_t high = size - 1u - nlow;         /* last cell of the spread region */

    if (cA == -1)
        cell[size - 1u - (uint32_t) __popc(mA & lt)] = lane;
    if (cB == -1)
        cell[size - 1u - nlowA - (uint32_t) __popc(mB & lt)] = lane + 32;
    if ((int) lane < nsym)
        next[lane] = (uint16_t) (cA == -1 ? 1 : cA);
    if ((int) lane + 32 < nsym)
        next[lane + 32] = (uint16_t) (cB == -1 ? 1 : cB);
    /* exclusive prefix sum of the positive counts, in symbol order */
    uint32_t pA = cA > 0 ? (uint32_t) cA : 0u, pB = cB > 0 ? (uint32_t) cB : 0u;
    uint32_t sA = pA, sB = pB;

#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        uint32_t tA = __shfl_up_sync(CRYO_FULL, sA, d), tB = __shfl_up_sync(CRYO_FULL, sB, d);

        if ((int) lane >= d)
        {
            sA += tA;
            sB += tB;
        }
    }
    uint32_t totA = __shfl_sync(CRYO_FULL, sA, 31);

    cum[lane] = (uint16_t) (sA - pA);
    cum[lane + 32] = (uint16_t) (totA + sB - pB);
    /* visit index of every low cell: position t is visited at step idx(t) = t * step^-1 */
    uint32_t inv = step;                            /* Newton: inv * step == 1 (mod 2^32) */

#pragma unroll
    for (int k = 0; k < 5; k++)
        inv *= 2u - step * inv;
    __syncwarp();
    /* spread: the j-th spread slot belongs to symbol s with cum[s] <= j < cum[s] + count[s];
     * it lands on the j-th visited position that is not a low cell */
    for (int s = 0; s < nsym; s++)
    {
        int c = counts[s];

        if (c <= 0)
            continue;
        uint32_t c0 = cum[s];

        for (uint32_t j = c0 + lane; j < c0 + (uint32_t) c; j += 32)
        {
            uint32_t i = j;

            if (nlow)
            {
                /* i = j + #{low cells visited at or before step i}: fixed point, monotone */
                for (;;)
                {
                    uint32_t k = 0;

                    for (uint32_t t = high + 1u; t < size; t++)
                        k += (((t * inv) & mask) <= i) ? 1u : 0u;
                    if (j + k == i)
                        break;
                    i = j + k;
                }
            }
            cell[(i * step) & mask] = (uint32_t) s;
        }
    }
    __syncwarp();
    /* state numbering in cell order: the r-th cell of symbol s gets next = count[s] + r */
    for (uint32_t p0 = 0; p0 < size; p0 += 32)
    {
        uint32_t p = p0 + lane;
        uint32_t s = cell[p];
        uint32_t m = zsw_match_any(s);
        uint32_t nx = (uint32_t) next[s] + (uint32_t) __popc(m & lt);

        __syncwarp();
        if ((m & lt) == 0)
            next[s] = (uint16_t) (next[s] + __popc(m));
        uint32_t nb = (uint32_t) (log - zs_highbit(nx));

        cell[p] = s | (nb << 8) | ((((nx << nb) - size) & 0xFFFFu) << 16);
        __syncwarp();
    }
}

/* warp-uniform decoder state that survives across the blocks of a frame */
struct ZswState
{
    int         huf_log, ll_log, of_log, ml_log;        /* -1 = undefined */
    const uint8_t *huf_desc;                            /* last Huffman tree description */
    uint32_t    huf_desc_len;
    uint32_t    rep0, rep1, rep2;
    /* how each sequence table was last defined, for Repeat_Mode: 0 predefined, 1 RLE, 2 FSE */
    int         tmode[3];
    const uint8_t *tdesc[3];
    uint32_t    tlen[3];
};

/* one sequence table (t: 0 LL, 1 OF, 2 ML); returns bytes of description consumed or ~0u */
CRYO_DEV uint32_t zsw_seq_table(ZswState &z, int mode, int t, const uint8_t *p, uint32_t left,
                                uint32_t *cell, uint8_t *work, const uint32_t *predef, int &logv,
                                uint32_t lane)
{
    const int max_log = t == 1 ? 8 : 9, max_sym = t == 0 ? 35 : t == 1 ? 31 : 52;
    int16_t  *counts = reinterpret_cast<int16_t *>(work + ZW_COUNTS);
    uint16_t *next = reinterpret_cast<uint16_t *>(work + ZW_NEXT);
    uint16_t *cum = reinterpret_cast<uint16_t *>(work + ZW_NEXT + 128);
    bool      repeat = false;

    if (mode == 3)
    {
        /* Repeat_Mode: the cells may have been overwritten since (the Huffman build scratch
         * overlays them), so the table is rebuilt from how it was last defined */
        if (logv < 0 || z.tmode[t] < 0)
            return ~0u;
        mode = z.tmode[t];
        p = z.tdesc[t];
        left = z.tlen[t];
        repeat = true;
    }
    uint32_t ret;

    switch (mode)
    {
        case 0:
        {
            const uint32_t n = t == 1 ? 32u : 64u, o = t == 0 ? 0u : t == 1 ? 64u : 96u;

            for (uint32_t i = lane; i < n; i += 32)
                cell[i] = predef[o + i];
            logv = t == 1 ? 5 : 6;
            z.tmode[t] = 0;
            ret = 0;
            break;
        }
        case 1:
            if (left < 1 || p[0] > max_sym)
                return ~0u;
            if (lane == 0)
                cell[0] = p[0];                       /* nbits 0, base 0 */
            logv = 0;
            z.tmode[t] = 1;
            z.tdesc[t] = p;
            z.tlen[t] = 1;
            ret = repeat ? 0u : 1u;
            break;
        default:
        {
            int32_t  nsym = 0, log = 0;
            uint32_t used = 0;

            if (lane == 0)
                used = fse_read_counts(p, left, max_log, max_sym, counts, &nsym, &log);
            used = __shfl_sync(CRYO_FULL, used, 0);
            nsym = __shfl_sync(CRYO_FULL, nsym, 0);
            log = __shfl_sync(CRYO_FULL, log, 0);
            if (used == 0)
                return ~0u;
            __syncwarp();
            fse_build_table_warp(cell, counts, nsym, log, next, cum, lane);
            logv = log;
            z.tmode[t] = 2;
            z.tdesc[t] = p;
            z.tlen[t] = used;
            ret = repeat ? 0u : used;
            break;
        }
    }
    /* the number of extra bits of every cell's code goes into bits 26..30 (base < 2^9 leaves
     * them free), so the sequence walk needs one lookup per table and state */
    __syncwarp();
    for (uint32_t i = lane; i < (1u << logv); i += 32)
    {
        const uint32_t c = cell[i], sym = c & 0xFFu;
        const uint32_t xb = t == 1 ? sym : (t == 0 ? CRYO_GLD(ZS_LL_PACK[sym]) : CRYO_GLD(ZS_ML_PACK[sym])) >> 24;

        cell[i] = (c & 0x03FFFFFFu) | (xb << 26);
    }
    __syncwarp();
    return ret;
}

/*
 * Huffman tree description -> decoding table huf[1 << log] (u16: symbol | nbits << 8), same
 * result as huf_build_table (zstd_decode.cuh) with the per-symbol ranking done by the whole
 * warp: 32 symbols per step, rank inside the step by __match_any_sync, running per-weight
 * counters in shared memory.  Returns bytes used by the description, 0 on error.
 */
CRYO_DEV bool zsw_huf_table(uint8_t *weights, uint32_t nw, uint16_t *huf, uint16_t *symstart,
                            uint32_t *rankc, int32_t *log_out, uint32_t lane);

CRYO_DEV uint32_t zsw_huf_build(const uint8_t *src, uint32_t n, uint16_t *huf, uint8_t *work,
                                int32_t *log_out, uint32_t lane)
{
    uint8_t  *weights = work + ZW_WEIGHTS;
    uint16_t *symstart = reinterpret_cast<uint16_t *>(work + ZW_SYMSTART);
    uint32_t *wfse = reinterpret_cast<uint32_t *>(work + ZW_WFSE);
    int16_t  *wcounts = reinterpret_cast<int16_t *>(work + ZW_WCOUNTS);
    uint16_t *wnext = reinterpret_cast<uint16_t *>(work + ZW_NEXT + 3 * 128);
    uint32_t *rankc = reinterpret_cast<uint32_t *>(work + ZW_COUNTS);      /* u32[16] counts, u32[16] starts */
    uint32_t  used = 0, nw = 0;
    int       bad = 0;

    if (n == 0)
        return 0;
    uint32_t h = src[0];

    if (h >= 128)
    {
        nw = h - 127;
        used = 1 + (nw + 1) / 2;
        if (used > n)
            return 0;
        for (uint32_t i = lane; i < nw; i += 32)
        {
            uint32_t b = src[1 + i / 2];

            weights[i] = (uint8_t) ((i & 1) ? (b & 15u) : (b >> 4));
        }
        __syncwarp();
    }
    else
    {
        used = 1 + h;
        if (used > n || h == 0)
            return 0;
        if (lane == 0)
        {
            int32_t  nsym = 0, flog = 0;
            uint32_t hdr = fse_read_counts(src + 1, h, 6, 12, wcounts, &nsym, &flog);

            if (hdr == 0 || hdr >= h)
                bad = 1;
            else
            {
                BitsBack bb;

                fse_build_table(wfse, wcounts, nsym, flog, wnext);
                if (!bb_init(bb, src + 1 + hdr, h - hdr))
                    bad = 1;
                else
                {
                    bb_refill(bb);
                    uint32_t s1 = bb_read(bb, (uint32_t) flog);
                    uint32_t s2 = bb_read(bb, (uint32_t) flog);

                    for (;;)
                    {
                        if (nw > 253)
                        {
                            bad = 1;
                            break;
                        }
                        uint32_t c1 = wfse[s1];

                        weights[nw++] = (uint8_t) c1;
                        bb_refill(bb);
                        s1 = (c1 >> 16) + bb_read(bb, (c1 >> 8) & 0xFFu);
                        if (bb.remaining < 0)
                        {
                            weights[nw++] = (uint8_t) wfse[s2];
                            break;
                        }
                        uint32_t c2 = wfse[s2];

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
            }
        }
        bad = __shfl_sync(CRYO_FULL, bad, 0);
        nw = __shfl_sync(CRYO_FULL, nw, 0);
        if (bad)
            return 0;
        __syncwarp();
    }
    return zsw_huf_table(weights, nw, huf, symstart, rankc, log_out, lane) ? used : 0u;
}

/*
 * Huffman weights[0, nw) (shared memory, room for one more) -> decoding table.  Whole warp;
 * symstart u16[256] and rankc u32[32] are shared scratch.  False on an invalid weight set.
 */
CRYO_DEV bool zsw_huf_table(uint8_t *weights, uint32_t nw, uint16_t *huf, uint16_t *symstart,
                            uint32_t *rankc, int32_t *log_out, uint32_t lane)
{
    /* sum of 2^(w-1), implied last weight */
    uint32_t sum = 0, over = 0;

    for (uint32_t i = lane; i < nw; i += 32)
    {
        uint32_t w = weights[i];

        if (w > 11)
            over = 1;
        else if (w)
            sum += 1u << (w - 1);
    }
    sum = __reduce_add_sync(CRYO_FULL, sum);
    over = __reduce_or_sync(CRYO_FULL, over);
    if (over || sum == 0)
        return false;
    const int log = zs_highbit(sum) + 1;

    if (log > 11)
        return false;
    const uint32_t left = (1u << log) - sum;

    if (left & (left - 1))
        return false;
    if (lane == 0)
        weights[nw] = (uint8_t) (zs_highbit(left) + 1);
    nw += 1;
    if (lane < 16)
        rankc[lane] = 0;
    __syncwarp();
    /* pass 1: symbols per weight */
    const uint32_t lt = (1u << lane) - 1u;

    for (uint32_t s0 = 0; s0 < nw; s0 += 32)
    {
        const uint32_t s = s0 + lane;
        const uint32_t w = s < nw ? weights[s] : 0u;
        const uint32_t m = zsw_match_any(w);

        if (w && (m & lt) == 0)
            rankc[w] += (uint32_t) __popc(m);
        __syncwarp();
    }
    /* first cell of every weight class: cells ordered by ascending weight, then symbol */
    if (lane == 0)
    {
        uint32_t a = 0;

        for (int r = 1; r <= log; r++)
        {
            uint32_t c = rankc[r];

            rankc[16 + r] = a;
            a += c << (r - 1);
        }
    }
    __syncwarp();
    /* pass 2: first cell of every symbol */
    for (uint32_t s0 = 0; s0 < nw; s0 += 32)
    {
        const uint32_t s = s0 + lane;
        const uint32_t w = s < nw ? weights[s] : 0u;
        const uint32_t m = zsw_match_any(w);

        if (w)
            symstart[s] = (uint16_t) (rankc[16 + w] + ((uint32_t) __popc(m & lt) << (w - 1)));
        __syncwarp();
        if (w && (m & lt) == 0)
            rankc[16 + w] += (uint32_t) __popc(m) << (w - 1);
        __syncwarp();
    }
    /* fill: long codes (few cells) one symbol per lane, short codes by the whole warp */
    for (uint32_t s0 = 0; s0 < nw; s0 += 32)
    {
        const uint32_t s = s0 + lane;
        const uint32_t w = s < nw ? weights[s] : 0u;
        const uint32_t len = w ? 1u << (w - 1) : 0u;
        const uint32_t st = w ? symstart[s] : 0u;
        const uint16_t ent = (uint16_t) (s | ((uint32_t) (log + 1 - (int) w) << 8));

        if (len && len <= 8)
            for (uint32_t i = 0; i < len; i++)
                huf[st + i] = ent;
        uint32_t big = __ballot_sync(CRYO_FULL, len > 8);

        while (big)
        {
            const int      k = __ffs((int) big) - 1;
            const uint32_t klen = __shfl_sync(CRYO_FULL, len, k), kst = __shfl_sync(CRYO_FULL, st, k);
            const uint32_t kent = __shfl_sync(CRYO_FULL, (uint32_t) ent, k);

            for (uint32_t i = lane; i < klen; i += 32)
                huf[kst + i] = (uint16_t) kent;
            big &= big - 1;
        }
    }
    __syncwarp();
    *log_out = log;
    return true;
}

/*
 * One Huffman stream, one lane: `count` symbols to dst; returns false on corruption.
 * The accumulator is an explicit (hi, lo) register pair with the next bit at bit 31 of hi: a
 * table index is one shift of hi (log <= 11), consuming a code is one funnel shift.  Two
 * symbols per refill check (2 x 11 <= 32), four symbols per 32-bit store.
 */
CRYO_DEV bool zsw_huf_stream(const uint16_t *huf, int log, const uint8_t *src, uint32_t n,
                             uint8_t *dst, uint32_t count)
{
    if (n == 0)
        return false;
    /* words of the stream, addressed from the aligned word that holds its first byte; the
     * bytes in front of the stream inside that word are never consumed by a valid stream and
     * a corrupt one fails the final bit count */
    const uint32_t head = (uint32_t) ((uintptr_t) src & 3u);
    const uint32_t *wb = reinterpret_cast<const uint32_t *>(src - head);
    const uint32_t nbytes = head + n;
    int32_t  wi = (int32_t) ((nbytes - 1u) >> 2);           /* word of the last byte */
    uint32_t w = wb[wi];
    const uint32_t keep = nbytes - 4u * (uint32_t) wi;       /* 1..4 valid low bytes */

    if (keep < 4)
        w &= (1u << (8u * keep)) - 1u;
    if (wi == 0 && head)
        w &= ~0u << (8u * head);
    if ((w >> (8u * (keep - 1u))) == 0)
        return false;                                       /* no end mark in the last byte */
    const int hb = zs_highbit(w);
    /* accumulator (hi:lo), next bit at bit 31 of hi; `avail` valid bits */
    uint32_t hi = hb ? w << (32 - hb) : 0u, lo = 0u;
    int32_t  avail = hb;
    /* bits of the stream not yet in the accumulator */
    int32_t  below = (int32_t) (8u * (4u * (uint32_t) wi - head));
    const int32_t total = below + hb;                       /* stream bits under the end mark */
    int32_t  used = 0;
    uint32_t nextw;

    wi--;
    nextw = wi >= 0 ? wb[wi] : 0u;
    const uint32_t sh = 32u - (uint32_t) log;
    uint32_t i = 0;

#ifdef CRYO_EMU
#define ZSW_SHR_C(x, s) ((s) >= 32 ? 0u : (x) >> (s))
#define ZSW_SHL_C(x, s) ((s) >= 32 ? 0u : (x) << (s))
#else
#define ZSW_SHR_C(x, s) __funnelshift_rc((x), 0u, (uint32_t) (s))
#define ZSW_SHL_C(x, s) __funnelshift_lc(0u, (x), (uint32_t) (s))
#endif
/* entering a new 128-byte line: ask for the line after next, so the dependent word loads of
 * this stream find their data in L1 instead of paying an L2 round trip each */
#ifdef CRYO_EMU
#define ZSW_HPREFETCH()
#else
#define ZSW_HPREFETCH()
#endif
#define ZSW_HREFILL()                                                        \
    if (avail <= 32)                                                         \
    {                                                                        \
        hi |= ZSW_SHR_C(nextw, avail);                                       \
        lo = ZSW_SHL_C(nextw, 32 - avail);                                   \
        avail += 32;                                                         \
        wi--;                                                                \
        nextw = wi >= 0 ? wb[wi] : 0u;                                       \
        ZSW_HPREFETCH();                                                     \
    }
#define ZSW_HDEC(sym)                                                        \
    {                                                                        \
        const uint32_t ent = huf[hi >> sh];                                  \
        const uint32_t nb = ent >> 8;                                        \
        sym = ent & 0xFFu;                                                   \
        hi = __funnelshift_l(lo, hi, nb);                                    \
        lo <<= nb;                                                           \
        avail -= (int32_t) nb;                                               \
        used += (int32_t) nb;                                                \
    }
    /* head: until dst + i is 4-byte aligned */
    while (i < count && ((uintptr_t) (dst + i) & 3u))
    {
        uint32_t s;

        ZSW_HREFILL();
        ZSW_HDEC(s);
        dst[i++] = (uint8_t) s;
    }
    {
        uint32_t *d4 = reinterpret_cast<uint32_t *>(dst + i);
        const uint32_t quads = (count - i) >> 2;

        for (uint32_t q = 0; q < quads; q++)
        {
            uint32_t s0, s1, s2, s3;

            ZSW_HREFILL();
            ZSW_HDEC(s0);
            ZSW_HDEC(s1);
            ZSW_HREFILL();
            ZSW_HDEC(s2);
            ZSW_HDEC(s3);
            d4[q] = s0 | (s1 << 8) | (s2 << 16) | (s3 << 24);
        }
        i += quads << 2;
    }
    while (i < count)
    {
        uint32_t s;

        ZSW_HREFILL();
        ZSW_HDEC(s);
        dst[i++] = (uint8_t) s;
    }
#undef ZSW_HREFILL
#undef ZSW_HPREFETCH
#undef ZSW_HDEC
#undef ZSW_SHR_C
#undef ZSW_SHL_C
    (void) below;
    /* every bit under the end mark consumed, no more, no less */
    return used == total;
}

/* Huffman literals of one block -> dst (global).  The table lives in the idle ring region. */
CRYO_DEV int zsw_huffman_literals(ZswState &z, int lit_type, const uint8_t *p, uint32_t left,
                                  uint32_t regen, uint32_t streams, uint8_t *dst, uint8_t *smem,
                                  uint32_t lane)
{
    uint16_t *huf = reinterpret_cast<uint16_t *>(smem + ZSW_OFF_HUF);
    uint8_t  *work = smem + ZSW_OFF_HUFWORK;

    if (lit_type == 2)
    {
        int32_t  log = 0;
        uint32_t used = zsw_huf_build(p, left, huf, work, &log, lane);

        if (used == 0)
            return ST_FORMAT;
        z.huf_log = log;
        z.huf_desc = p;
        z.huf_desc_len = used;
        p += used;
        left -= used;
    }
    else
    {
        /* treeless: the table of the previous Huffman block; the ring has overwritten it
         * since, so rebuild it from the remembered tree description */
        int32_t log = 0;

        if (z.huf_log < 1 || zsw_huf_build(z.huf_desc, z.huf_desc_len, huf, work, &log, lane) == 0)
            return ST_FORMAT;
    }
    __syncwarp();
    bool ok = true;

    if (streams == 1)
    {
        if (lane == 0)
            ok = zsw_huf_stream(huf, z.huf_log, p, left, dst, regen);
    }
    else
    {
        if (left < 6)
            ok = false;
        else
        {
            uint32_t s1 = p[0] | ((uint32_t) p[1] << 8);
            uint32_t s2 = p[2] | ((uint32_t) p[3] << 8);
            uint32_t s3 = p[4] | ((uint32_t) p[5] << 8);
            uint32_t seg = (regen + 3) / 4;

            if (6 + s1 + s2 + s3 > left || seg * 3 > regen)
                ok = false;
            else if (lane < 4)
            {
                uint32_t s4 = left - 6 - s1 - s2 - s3;
                uint32_t so = lane == 0 ? 0 : lane == 1 ? s1 : lane == 2 ? s1 + s2 : s1 + s2 + s3;
                uint32_t sn = lane == 0 ? s1 : lane == 1 ? s2 : lane == 2 ? s3 : s4;
                uint32_t cnt = lane < 3 ? seg : regen - 3 * seg;

                ok = zsw_huf_stream(huf, z.huf_log, p + 6 + so, sn, dst + lane * seg, cnt);
            }
        }
    }
    __threadfence_block();
    return __any_sync(CRYO_FULL, !ok) ? ST_FORMAT : ST_OK;
}

/* literal source of one block for the sequence phase */
struct ZswLits
{
    const uint8_t *abase;       /* 16-byte aligned address at or before the first literal */
    uint8_t    *win;            /* shared window */
    uint32_t    delta, n, pos, wbase, lim;
    bool        rle, wvalid;
    uint8_t     rle_byte;
};

CRYO_DEV void zsw_lits_fill(ZswLits &L, uint32_t ip, uint32_t lane)
{
    __syncwarp();
    L.wbase = ip & ~15u;
    L.wvalid = true;
    {
        uint32_t a = L.wbase + 16u * lane;

        if (a < L.lim)
            st16(L.win + 16u * lane, ld16(L.abase + a));
    }
    __syncwarp();
}

CRYO_DEV void zsw_lits_emit(WOut &o, ZswLits &L, uint32_t n, uint32_t lane)
{
    if (n == 0)
        return;
    if (L.rle)
        wx_fill_byte(o, L.rle_byte, n, lane);
    else if (n >= WX_BULK || n + 16u > ZSW_LITWIN)
    {
        /* long run: straight from global memory (through the ring when it is short of a bulk) */
        if (n >= WX_BULK)
            wx_literals(o, L.abase + L.delta + L.pos, n, lane);
        else
        {
            const uint8_t *g = L.abase + L.delta + L.pos;

            for (uint32_t i = lane; i < n; i += 32)
                o.ring[(o.pos + i) & WX_RMASK] = g[i];
            o.pos += n;
            __syncwarp();
            wx_drain(o, lane);
        }
    }
    else
    {
        uint32_t ip = L.delta + L.pos;

        if (!L.wvalid || ip + n > L.wbase + ZSW_LITWIN)
            zsw_lits_fill(L, ip, lane);
        wx_literals(o, L.win + (ip - L.wbase), n, lane);
    }
    L.pos += n;
}

/*
 * Backward bit reader over a shared-memory window of the sequence bitstream.  Positions are
 * bit offsets from `abase` (16-byte aligned, at or before the stream); bits [lowbit, bitpos)
 * are unread.  A read takes the 64 bits below bitpos out of three aligned words; there is no
 * accumulator to maintain.  The window slides down as the stream is consumed; below the
 * stream start it reads zeros (the 16 bytes in front of the window are kept zero for that).
 */
struct ZswBits
{
    const uint8_t *abase;
    const uint32_t *w32;        /* shared: word 0 = abase[wlo .. wlo + 4) */
    uint8_t    *win;            /* shared: 16 zero bytes, then the window */
    uint32_t    wlo;            /* byte offset of the window in abase coordinates, multiple of 16 */
    uint32_t    lim;            /* stream end rounded up to 16 (abase coordinates) */
    uint32_t    bitpos, lowbit;
};

CRYO_DEV void zsw_bits_fill(ZswBits &B, uint32_t lane)
{
    const uint32_t topbyte = B.bitpos >> 3;     /* highest byte still needed */

    __syncwarp();
    B.wlo = topbyte + 16u > ZSW_SEQWIN ? ((topbyte + 16u - ZSW_SEQWIN) & ~15u) : 0u;
    {
        const uint32_t a = B.wlo + 16u * lane;
        uint4 v = make_uint4(0, 0, 0, 0);

        if (a < B.lim)
            v = ld16(B.abase + a);
        st16(B.win + 16u + 16u * lane, v);
    }
    __syncwarp();
    if (B.wlo == 0 && lane < (B.lowbit >> 3))
        B.win[16u + lane] = 0;                  /* bytes in front of the stream read as zero */
    __syncwarp();
}

CRYO_DEV bool zsw_bits_init(ZswBits &B, const uint8_t *p, uint32_t n, uint8_t *win, uint32_t lane)
{
    if (n == 0)
        return false;
    const uint32_t last = p[n - 1];

    if (last == 0)
        return false;
    const uint32_t delta = (uint32_t) ((uintptr_t) p & 15u);

    B.abase = p - delta;
    B.win = win;
    B.w32 = reinterpret_cast<const uint32_t *>(win + 16);
    B.lowbit = delta * 8u;
    B.bitpos = (delta + n - 1u) * 8u + (uint32_t) zs_highbit(last);
    B.lim = (delta + n + 15u) & ~15u;
    if (lane < 4)
        reinterpret_cast<uint32_t *>(win)[lane] = 0;
    zsw_bits_fill(B, lane);
    return true;
}

/* the 64 bits below bitpos as (hi, lo); callers keep bitpos - 64 - 32 >= wlo * 8 or wlo == 0 */
CRYO_DEV void zsw_bits_peek(const ZswBits &B, uint32_t &hi, uint32_t &lo)
{
    const int32_t rel = (int32_t) B.bitpos - 64 - (int32_t) (B.wlo * 8u);
    const int32_t wi = rel >> 5;
    const uint32_t sh = (uint32_t) rel & 31u;
    const uint32_t w0 = B.w32[wi], w1 = B.w32[wi + 1], w2 = B.w32[wi + 2];

    lo = __funnelshift_r(w0, w1, sh);
    hi = __funnelshift_r(w1, w2, sh);
}

/* n bits (n <= 32) that follow the first c bits of the peeked window, c + n <= 64 */
CRYO_DEV uint32_t zsw_bits_get(uint32_t hi, uint32_t lo, uint32_t c, uint32_t n)
{
    const uint32_t top = c < 32u ? __funnelshift_l(lo, hi, c) : (lo << (c - 32u));

    return n ? top >> (32u - n) : 0u;
}

/*
 * Decode the zstd frame(s) at src[0, csize) into out[0, cap).  One warp; `smem` is this
 * warp's ZSW_PER_WARP bytes; `scratch` is ZSTDD_SCRATCH_BYTES of global memory private to
 * this warp (16-byte aligned); predef holds the three predefined FSE tables.
 */
CRYO_DEV void zstdw_decode_frame(const uint8_t *src, uint32_t csize, uint8_t *out, uint32_t cap,
                                 uint32_t *out_size, int32_t *status, uint8_t *scratch,
                                 const uint32_t *predef, uint8_t *smem, uint32_t lane)
{
    WOut     o;
    ZswState z;
    int      err = ST_OK;
    uint32_t ip = 0;
    const uint8_t *in = src;
    uint32_t *ll_tab = reinterpret_cast<uint32_t *>(smem + ZSW_OFF_LL);
    uint32_t *of_tab = reinterpret_cast<uint32_t *>(smem + ZSW_OFF_OF);
    uint32_t *ml_tab = reinterpret_cast<uint32_t *>(smem + ZSW_OFF_ML);

    wx_init(o, out, cap, smem + ZSW_OFF_RING);
#ifndef CRYO_EMU
    /* pull the compressed frame into L2 now: every later read of it is on a dependent chain */
    for (uint32_t a = 128u * lane; a < csize; a += 128u * 32u)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(src + a));
#endif

    while (err == ST_OK && ip < csize)
    {
        /* ---- frame header ---- */
        if (ip + 4 > csize)
        {
            err = ST_INPUT;
            break;
        }
        uint32_t magic = in[ip] | ((uint32_t) in[ip + 1] << 8) | ((uint32_t) in[ip + 2] << 16) |
                         ((uint32_t) in[ip + 3] << 24);

        if ((magic & 0xFFFFFFF0u) == 0x184D2A50u)
        {
            if (ip + 8 > csize)
            {
                err = ST_INPUT;
                break;
            }
            uint32_t len = in[ip + 4] | ((uint32_t) in[ip + 5] << 8) | ((uint32_t) in[ip + 6] << 16) |
                           ((uint32_t) in[ip + 7] << 24);

            if (len > csize - ip - 8)
            {
                err = ST_INPUT;
                break;
            }
            ip += 8 + len;
            continue;
        }
        if (magic != 0xFD2FB528u)
        {
            err = ST_FORMAT;
            break;
        }
        ip += 4;
        if (ip + 1 > csize)
        {
            err = ST_INPUT;
            break;
        }
        uint32_t fhd = in[ip++];
        uint32_t fcs_flag = fhd >> 6, single = (fhd >> 5) & 1u, checksum = (fhd >> 2) & 1u;
        uint32_t dict_flag = fhd & 3u;
        uint64_t fcs = 0;

        if (fhd & 0x08u)
        {
            err = ST_FORMAT;
            break;
        }
        if (!single)
        {
            if (ip + 1 > csize)
            {
                err = ST_INPUT;
                break;
            }
            uint32_t b = in[ip++];

            if (10 + (b >> 3) > 27)
            {
                err = ST_FORMAT;        /* ZSTD_decompress' default window limit */
                break;
            }
        }
        uint32_t dict_bytes = dict_flag == 3 ? 4 : dict_flag;
        uint32_t fcs_bytes = fcs_flag == 0 ? (single ? 1u : 0u) : (1u << fcs_flag);

        if (ip + dict_bytes + fcs_bytes > csize)
        {
            err = ST_INPUT;
            break;
        }
        uint32_t dict_id = 0;

        for (uint32_t i = 0; i < dict_bytes; i++)
            dict_id |= (uint32_t) in[ip + i] << (8 * i);
        ip += dict_bytes;
        if (dict_id != 0)
        {
            err = ST_FORMAT;            /* no dictionary on this path */
            break;
        }
        for (uint32_t i = 0; i < fcs_bytes; i++)
            fcs |= (uint64_t) in[ip + i] << (8 * i);
        if (fcs_bytes == 2)
            fcs += 256;
        ip += fcs_bytes;
        /* RFC 8878 says min(Window_Size, 128 KiB); libzstd 1.5.5's ZSTD_decompress (the
         * reference's call, compression.c:116) only enforces the constant -- follow it */
        const uint32_t block_max = ZS_MAXBLOCK;
        const uint32_t frame_start = o.pos;

        z.huf_log = z.ll_log = z.of_log = z.ml_log = -1;
        z.huf_desc = nullptr;
        z.huf_desc_len = 0;
        z.rep0 = 1;
        z.rep1 = 4;
        z.rep2 = 8;
        z.tmode[0] = z.tmode[1] = z.tmode[2] = -1;
        z.tdesc[0] = z.tdesc[1] = z.tdesc[2] = nullptr;
        z.tlen[0] = z.tlen[1] = z.tlen[2] = 0;

        /* ---- blocks ---- */
        for (;;)
        {
            if (ip + 3 > csize)
            {
                err = ST_INPUT;
                break;
            }
            uint32_t bh = in[ip] | ((uint32_t) in[ip + 1] << 8) | ((uint32_t) in[ip + 2] << 16);
            uint32_t last = bh & 1u, type = (bh >> 1) & 3u, bsize = bh >> 3;

            ip += 3;
            if (type == 3 || bsize > block_max)
            {
                err = ST_FORMAT;
                break;
            }
            if (type == 0)
            {
                if (bsize > csize - ip)
                {
                    err = ST_INPUT;
                    break;
                }
                if (bsize > cap - o.pos)
                {
                    err = ST_OUTPUT;
                    break;
                }
                if (bsize)
                    wx_literals(o, in + ip, bsize, lane);
                ip += bsize;
            }
            else if (type == 1)
            {
                if (ip + 1 > csize)
                {
                    err = ST_INPUT;
                    break;
                }
                if (bsize > cap - o.pos)
                {
                    err = ST_OUTPUT;
                    break;
                }
                if (bsize)
                    wx_fill_byte(o, in[ip], bsize, lane);
                ip += 1;
            }
            else
            {
                /* ---- Compressed block ---- */
                if (bsize == 0 || bsize > csize - ip)
                {
                    err = bsize == 0 ? ST_FORMAT : ST_INPUT;
                    break;
                }
                const uint8_t *bp = in + ip;
                const uint32_t block_start = o.pos;
                uint32_t lt = bp[0] & 3u, sf = (bp[0] >> 2) & 3u;
                uint32_t lhdr, regen, lcsize = 0, streams = 1;
                const uint8_t *lit_base;
                ZswLits  L;

                L.rle = false;
                L.rle_byte = 0;
                if (lt < 2)
                {
                    if (sf == 0 || sf == 2)
                    {
                        lhdr = 1;
                        regen = bp[0] >> 3;
                    }
                    else if (sf == 1)
                    {
                        lhdr = 2;
                        regen = bsize >= 2 ? ((bp[0] >> 4) | ((uint32_t) bp[1] << 4)) : 0;
                    }
                    else
                    {
                        lhdr = 3;
                        regen = bsize >= 3 ? ((bp[0] >> 4) | ((uint32_t) bp[1] << 4) |
                                              ((uint32_t) bp[2] << 12)) : 0;
                    }
                    lcsize = lt == 0 ? regen : 1;
                    if (lhdr + lcsize > bsize || regen > ZS_MAXBLOCK)
                    {
                        err = ST_FORMAT;
                        break;
                    }
                    lit_base = bp + lhdr;
                    if (lt == 1)
                    {
                        L.rle = true;
                        L.rle_byte = bp[lhdr];
                    }
                }
                else
                {
                    if (bsize < 5)
                    {
                        err = ST_FORMAT;
                        break;
                    }
                    uint64_t v = bp[0] | ((uint64_t) bp[1] << 8) | ((uint64_t) bp[2] << 16) |
                                 ((uint64_t) bp[3] << 24) | ((uint64_t) bp[4] << 32);

                    if (sf < 2)
                    {
                        lhdr = 3;
                        regen = (uint32_t) (v >> 4) & 0x3FFu;
                        lcsize = (uint32_t) (v >> 14) & 0x3FFu;
                        streams = sf == 0 ? 1 : 4;
                    }
                    else if (sf == 2)
                    {
                        lhdr = 4;
                        regen = (uint32_t) (v >> 4) & 0x3FFFu;
                        lcsize = (uint32_t) (v >> 18) & 0x3FFFu;
                        streams = 4;
                    }
                    else
                    {
                        lhdr = 5;
                        regen = (uint32_t) (v >> 4) & 0x3FFFFu;
                        lcsize = (uint32_t) (v >> 22) & 0x3FFFFu;
                        streams = 4;
                    }
                    if (lhdr + lcsize > bsize || regen > ZS_MAXBLOCK)
                    {
                        err = ST_FORMAT;
                        break;
                    }
                    lit_base = scratch;
                }
                /* sequences header */
                uint32_t sp = lhdr + lcsize, nseq;

                if (sp + 1 > bsize)
                {
                    err = ST_INPUT;
                    break;
                }
                if (bp[sp] < 128)
                {
                    nseq = bp[sp];
                    sp += 1;
                }
                else if (bp[sp] < 255)
                {
                    if (sp + 2 > bsize)
                    {
                        err = ST_INPUT;
                        break;
                    }
                    nseq = ((uint32_t) (bp[sp] - 128) << 8) + bp[sp + 1];
                    sp += 2;
                }
                else
                {
                    if (sp + 3 > bsize)
                    {
                        err = ST_INPUT;
                        break;
                    }
                    nseq = bp[sp + 1] + ((uint32_t) bp[sp + 2] << 8) + 0x7F00u;
                    sp += 3;
                }
                int modes = 0;

                if (nseq)
                {
                    if (sp + 1 > bsize)
                    {
                        err = ST_INPUT;
                        break;
                    }
                    modes = bp[sp++];
                    if (modes & 3)
                    {
                        err = ST_FORMAT;
                        break;
                    }
                }
                else if (sp != bsize)
                {
                    err = ST_INPUT;
                    break;
                }
                /* ---- entropy phase: the ring's shared memory holds tables and scratch ---- */
                if (lt >= 2)
                    wx_drain_all(o, lane);          /* the Huffman table overlays the ring */
                if (lt >= 2)
                {
                    err = zsw_huffman_literals(z, (int) lt, bp + lhdr, lcsize, regen, streams,
                                               scratch, smem, lane);
                    if (err != ST_OK)
                        break;
                }
                if (nseq)
                {
                    uint8_t *work = smem + ZSW_OFF_FSEWORK;
                    uint32_t u;

                    __syncwarp();
                    u = zsw_seq_table(z, (modes >> 6) & 3, 0, bp + sp, bsize - sp, ll_tab, work, predef,
                                      z.ll_log, lane);
                    if (u == ~0u)
                    {
                        err = ST_FORMAT;
                        break;
                    }
                    sp += u;
                    u = zsw_seq_table(z, (modes >> 4) & 3, 1, bp + sp, bsize - sp, of_tab, work, predef,
                                      z.of_log, lane);
                    if (u == ~0u)
                    {
                        err = ST_FORMAT;
                        break;
                    }
                    sp += u;
                    u = zsw_seq_table(z, (modes >> 2) & 3, 2, bp + sp, bsize - sp, ml_tab, work, predef,
                                      z.ml_log, lane);
                    if (u == ~0u)
                    {
                        err = ST_FORMAT;
                        break;
                    }
                    sp += u;
                }
                if (lt >= 2)
                    wx_after_bulk(o, 0, lane);      /* the ring is garbage now: re-prime it */

                L.abase = lit_base - ((uintptr_t) lit_base & 15u);
                L.delta = (uint32_t) ((uintptr_t) lit_base & 15u);
                L.n = regen;
                L.pos = 0;
                L.win = smem + ZSW_OFF_LITWIN;
                L.wvalid = false;
                L.wbase = 0;
                L.lim = (L.delta + regen + 15u) & ~15u;

                /* ---- sequence phase ---- */
                if (nseq)
                {
                    ZswBits  B;
                    const uint32_t ll_log = (uint32_t) z.ll_log, of_log = (uint32_t) z.of_log,
                                   ml_log = (uint32_t) z.ml_log;

                    if (sp > bsize || !zsw_bits_init(B, bp + sp, bsize - sp, smem + ZSW_OFF_SEQWIN, lane))
                    {
                        err = ST_FORMAT;
                        break;
                    }
                    uint32_t sl, so, sm;
                    {
                        uint32_t hi, lo;

                        zsw_bits_peek(B, hi, lo);
                        sl = zsw_bits_get(hi, lo, 0, ll_log);
                        so = zsw_bits_get(hi, lo, ll_log, of_log);
                        sm = zsw_bits_get(hi, lo, ll_log + of_log, ml_log);
                        if (B.bitpos - B.lowbit < ll_log + of_log + ml_log)
                        {
                            err = ST_INPUT;
                            break;
                        }
                        B.bitpos -= ll_log + of_log + ml_log;
                    }
                    uint32_t rep0 = z.rep0, rep1 = z.rep1, rep2 = z.rep2;
                    uint32_t lpos = 0;                  /* literals consumed (mirrors L.pos) */

                    /*
                     * 32 sequences at a time.  Pass 1 (warp-uniform, serial): walk the three FSE
                     * states; only the state bits are read here, lane k keeps the cells and the
                     * bit position of sequence k.  Pass 2 (one sequence per lane): every lane
                     * extracts its own offset / match-length / literal-length extra bits.  Then
                     * the 32 sequences are executed in order (repeat offsets, checks, copies).
                     */
                    for (uint32_t done = 0; done < nseq && err == ST_OK; done += 32)
                    {
                        const uint32_t g = nseq - done < 32u ? nseq - done : 32u;
                        uint32_t my_cl = 0, my_co = 0, my_cm = 0, my_bp = 0;
                        int32_t  under = 0;

                        /* 32 sequences take at most 32 x 89 bits; keep them and a peek inside the window */
                        if (B.wlo != 0 && B.bitpos < B.wlo * 8u + 3200u)
                            zsw_bits_fill(B, lane);
                        const int32_t wbits = (int32_t) (B.wlo * 8u);

                        for (uint32_t k = 0; k < g; k++)
                        {
                            const uint32_t cl = ll_tab[sl], co = of_tab[so], cm = ml_tab[sm];
                            const uint32_t text = (cl >> 26) + (co >> 26) + (cm >> 26);

                            if (lane == k)
                            {
                                my_cl = cl;
                                my_co = co;
                                my_cm = cm;
                                my_bp = B.bitpos;
                            }
                            if (done + k + 1 < nseq)
                            {
                                const uint32_t nbl = (cl >> 8) & 0xFFu, nbm = (cm >> 8) & 0xFFu,
                                               nbo = (co >> 8) & 0xFFu;
                                const uint32_t p = B.bitpos - text;         /* state bits end here */
                                const int32_t  rel = (int32_t) p - 32 - wbits;
                                const int32_t  wi = rel >> 5;
                                uint32_t top = __funnelshift_r(B.w32[wi], B.w32[wi + 1], (uint32_t) rel & 31u);

                                sl = ((cl >> 16) & 0x3FFu) + __funnelshift_l(top, 0u, nbl);
                                top <<= nbl;
                                sm = ((cm >> 16) & 0x3FFu) + __funnelshift_l(top, 0u, nbm);
                                top <<= nbm;
                                so = ((co >> 16) & 0x3FFu) + __funnelshift_l(top, 0u, nbo);
                                B.bitpos = p - (nbl + nbm + nbo);
                            }
                            else
                                B.bitpos -= text;
                            under |= (int32_t) (B.bitpos - B.lowbit);
                        }
                        if (under < 0)
                        {
                            err = ST_INPUT;         /* the stream ended inside a sequence */
                            break;
                        }
                        /* pass 2: lane k = sequence done + k */
                        uint32_t my_ov = 0, my_ml = 0, my_ll = 0;

                        if (lane < g)
                        {
                            const uint32_t xo = my_co >> 26, xm = my_cm >> 26, xl = my_cl >> 26;
                            const int32_t  rel = (int32_t) my_bp - 64 - wbits;
                            const int32_t  wi = rel >> 5;
                            const uint32_t sh = (uint32_t) rel & 31u;
                            const uint32_t w0 = B.w32[wi], w1 = B.w32[wi + 1], w2 = B.w32[wi + 2];
                            const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);

                            my_ov = (1u << xo) + zsw_bits_get(hi, lo, 0, xo);
                            my_ml = (CRYO_GLD(ZS_ML_PACK[my_cm & 0xFFu]) & 0xFFFFFFu) + zsw_bits_get(hi, lo, xo, xm);
                            my_ll = (CRYO_GLD(ZS_LL_PACK[my_cl & 0xFFu]) & 0xFFFFFFu) + zsw_bits_get(hi, lo, xo + xm, xl);
                        }
                        /* execution, in order */
                        for (uint32_t k = 0; k < g; k++)
                        {
                            const uint32_t ov = __shfl_sync(CRYO_FULL, my_ov, (int) k);
                            const uint32_t ml = __shfl_sync(CRYO_FULL, my_ml, (int) k);
                            const uint32_t ll = __shfl_sync(CRYO_FULL, my_ll, (int) k);
                            uint32_t off;

                            if (ov > 3)
                            {
                                off = ov - 3;
                                rep2 = rep1;
                                rep1 = rep0;
                                rep0 = off;
                            }
                            else
                            {
                                const uint32_t idx = ov - 1 + (ll == 0 ? 1u : 0u);

                                if (idx == 0)
                                    off = rep0;
                                else
                                {
                                    off = idx == 1 ? rep1 : idx == 2 ? rep2 : rep0 - 1;
                                    if (idx > 1)
                                        rep2 = rep1;
                                    rep1 = rep0;
                                    rep0 = off;
                                }
                            }
                            const uint32_t mpos = o.pos + ll, epos = mpos + ml;     /* < 2^28: no wrap */

                            if ((ll > regen - lpos) | (epos > cap) | (epos - block_start > block_max) |
                                (off - 1u >= mpos - frame_start))
                            {
                                err = ll > regen - lpos ? ST_FORMAT
                                      : epos > cap ? ST_OUTPUT
                                      : epos - block_start > block_max ? ST_FORMAT : ST_OFFSET;
                                break;
                            }
                            /*
                             * Fast path: a literal run of up to 64 bytes served from the literal
                             * window and a non-overlapping match of up to 64 bytes whose source is
                             * in the ring: two predicated shared-memory moves each.
                             */
                            const uint32_t lip = L.delta + lpos;

                            if (ll <= 64u && ml <= 64u && !L.rle && off >= ml && off <= WX_RING - 64u &&
                                mpos - off >= o.lo)
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
                                    o.ring[(mpos + lane) & WX_RMASK] = o.ring[(mpos - off + lane) & WX_RMASK];
                                if (lane + 32u < ml)
                                    o.ring[(mpos + lane + 32u) & WX_RMASK] =
                                        o.ring[(mpos - off + lane + 32u) & WX_RMASK];
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
                            wx_match(o, off, ml, lane);
                        }
                    }
                    z.rep0 = rep0;
                    z.rep1 = rep1;
                    z.rep2 = rep2;
                    L.pos = lpos;
                    if (err != ST_OK)
                        break;
                    if (B.bitpos != B.lowbit)
                    {
                        err = ST_INPUT;
                        break;
                    }
                }
                /* literals left after the last sequence */
                uint32_t rest = L.n - L.pos;

                if (rest > cap - o.pos)
                {
                    err = ST_OUTPUT;
                    break;
                }
                if (o.pos + rest - block_start > block_max)
                {
                    err = ST_FORMAT;
                    break;
                }
                zsw_lits_emit(o, L, rest, lane);
                ip += bsize;
            }
            if (last)
                break;
        }
        if (err != ST_OK)
            break;
        if (fcs_bytes && (uint64_t) (o.pos - frame_start) != fcs)
        {
            err = ST_SIZE;
            break;
        }
        if (checksum)
        {
            if (ip + 4 > csize)
            {
                err = ST_INPUT;
                break;
            }
            ip += 4;                    /* XXH64 content checksum: skipped, not verified */
        }
    }
    wx_drain_all(o, lane);
    if (lane == 0)
    {
        *out_size = err == ST_OK ? o.pos : 0u;
        *status = err;
    }
}

/* build the three predefined tables into predef[ZSW_PREDEF_CELLS] (one warp, once per context) */
CRYO_DEV void zsw_build_predef(uint32_t *predef, uint8_t *smem, uint32_t lane)
{
    int16_t  *counts = reinterpret_cast<int16_t *>(smem);
    uint16_t *next = reinterpret_cast<uint16_t *>(smem + 256);
    uint16_t *cum = reinterpret_cast<uint16_t *>(smem + 512);
    uint32_t *cell = reinterpret_cast<uint32_t *>(smem + 1024);

    for (int t = 0; t < 3; t++)
    {
        const int n = t == 0 ? 36 : t == 1 ? 29 : 53, log = t == 1 ? 5 : 6;
        const int16_t *def = t == 0 ? ZS_LL_DEFAULT : t == 1 ? ZS_OF_DEFAULT : ZS_ML_DEFAULT;
        const uint32_t o = t == 0 ? 0u : t == 1 ? 64u : 96u;

        for (int i = (int) lane; i < n; i += 32)
            counts[i] = def[i];
        __syncwarp();
        fse_build_table_warp(cell, counts, n, log, next, cum, lane);
        __syncwarp();
        for (uint32_t i = lane; i < (1u << log); i += 32)
            predef[o + i] = cell[i];
        __syncwarp();
    }
}
