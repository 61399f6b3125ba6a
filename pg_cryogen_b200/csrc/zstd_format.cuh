/*
 * zstd_format.cuh -- what every zstd kernel shares (RFC 8878): the code tables of the sequence
 * section, the backward bit reader, FSE table descriptions and decoding tables, and the Huffman
 * tree description -> decoding table step.  Decoders: zstd_decode_w.cuh (one warp per frame),
 * zstd_decode_p.cuh (the phase-split pipeline) with zstd_decode_c.cuh (its CTA-per-frame stage 4);
 * the encoder (zstd_encode.cuh) uses the table builder.
 */
#pragma once
#include "cryo_common.cuh"

#define ZS_MAXBLOCK   (1u << 17)
#define ZSTDD_SCRATCH_BYTES (ZS_MAXBLOCK + 256)


/* work-area offsets (bytes from ZS_SM_WORK) */
#define ZW_WEIGHTS    0       /* u8[256]  Huffman weights */
#define ZW_SYMSTART   256     /* u16[256] first table cell of every symbol */
#define ZW_WFSE       768     /* u32[64]  FSE table of the Huffman weights */
#define ZW_WCOUNTS    1024    /* i16[16] */
#define ZW_COUNTS     1152    /* i16[64] x 3: LL, OF, ML normalised counts */
#define ZW_NEXT       1536    /* u16[64] x 4: per-builder scratch */


#ifdef CRYO_EMU
#define CRYO_CONST static const
#else
#define CRYO_CONST __constant__
#endif

CRYO_CONST int16_t ZS_LL_DEFAULT[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2,
    2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
CRYO_CONST int16_t ZS_ML_DEFAULT[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1,
    1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1,
    -1, -1, -1, -1, -1};
CRYO_CONST int16_t ZS_OF_DEFAULT[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1,
    1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};
CRYO_CONST uint32_t ZS_LL_BASE[36] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16,
    18, 20, 22, 24, 28, 32, 40, 48, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768,
    65536};
CRYO_CONST uint8_t ZS_LL_BITS[36] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1,
    1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
CRYO_CONST uint32_t ZS_ML_BASE[53] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18,
    19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 37, 39, 41, 43, 47, 51,
    59, 67, 83, 99, 131, 259, 515, 1027, 2051, 4099, 8195, 16387, 32771, 65539};
CRYO_CONST uint8_t ZS_ML_BITS[53] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11,
    12, 13, 14, 15, 16};

CRYO_DEV int zs_highbit(uint32_t v) { return 31 - __clz((int) v); }

/* ---- backward bitstream (Huffman streams, FSE sequences) ----------------- */

struct BitsBack
{
    uintptr_t   start;          /* address of the first stream byte */
    uintptr_t   cur;            /* aligned address of the word held in nextw */
    uint64_t    acc;            /* next bit to read is bit 63 */
    uint32_t    nextw;
    int32_t     avail;          /* valid bits in acc */
    int32_t     remaining;      /* stream bits not yet consumed; < 0 = read past the start */
};

CRYO_DEV uint32_t bb_load(uintptr_t addr, uintptr_t start)
{
    if (addr + 4 <= start)
        return 0u;
    uint32_t w = *reinterpret_cast<const uint32_t *>(addr);

    if (addr < start)
        w &= ~0u << (8u * (uint32_t) (start - addr));
    return w;
}

/* returns false when the stream is empty or its last byte is zero (no end marker) */
CRYO_DEV bool bb_init(BitsBack &b, const uint8_t *p, uint32_t n)
{
    if (n == 0)
        return false;
    uintptr_t A = (uintptr_t) p, E = A + n;
    uintptr_t wl = (E - 1) & ~(uintptr_t) 3;
    uint32_t  w = *reinterpret_cast<const uint32_t *>(wl);
    uint32_t  keep = (uint32_t) (E - wl);           /* 1..4 valid low bytes */

    if (keep < 4)
        w &= (1u << (8u * keep)) - 1u;
    if (wl < A)
        w &= ~0u << (8u * (uint32_t) (A - wl));
    if ((w >> (8u * (keep - 1u))) == 0)
        return false;
    int hb = zs_highbit(w);

    b.start = A;
    b.acc = hb ? ((uint64_t) w << (64 - hb)) : 0ull;
    b.avail = hb;
    b.remaining = (int32_t) ((n - 1) * 8u) + (hb - 8 * (int) (keep - 1));
    b.cur = wl - 4;
    b.nextw = bb_load(b.cur, A);
    return true;
}

CRYO_DEV void bb_refill(BitsBack &b)
{
    if (b.avail <= 32)
    {
        b.acc |= (uint64_t) b.nextw << (32 - b.avail);
        b.avail += 32;
        b.cur -= 4;
        b.nextw = bb_load(b.cur, b.start);
    }
}

/* nb <= 32 and nb <= avail (callers refill first) */
CRYO_DEV uint32_t bb_read(BitsBack &b, uint32_t nb)
{
    uint32_t v = nb ? (uint32_t) (b.acc >> (64 - nb)) : 0u;

    b.acc = nb ? (b.acc << nb) : b.acc;
    b.avail -= (int32_t) nb;
    b.remaining -= (int32_t) nb;
    return v;
}

/* ---- forward bit reader for FSE table descriptions (single lane) --------- */

CRYO_DEV uint32_t fw_peek(const uint8_t *p, uint32_t n, uint32_t bit, uint32_t k)
{
    uint32_t b = bit >> 3, v = 0;

#pragma unroll
    for (uint32_t i = 0; i < 4; i++)
        if (b + i < n)
            v |= (uint32_t) p[b + i] << (8u * i);
    return (v >> (bit & 7u)) & ((1u << k) - 1u);
}

/*
 * Normalised counts (RFC 8878 4.1.1).  Single lane.  Returns bytes used, 0 on error.
 */
CRYO_DEV uint32_t fse_read_counts(const uint8_t *p, uint32_t n, int max_log, int max_sym,
                                  int16_t *counts, int32_t *nsym_out, int32_t *log_out)
{
    if (n == 0)
        return 0;
    int      log = (int) fw_peek(p, n, 0, 4) + 5;
    uint32_t bit = 4;
    int      remaining, sym = 0;

    if (log > max_log)
        return 0;
    remaining = 1 << log;
    while (remaining > 0 && sym <= max_sym)
    {
        int      bits = zs_highbit((uint32_t) remaining + 1u) + 1;
        uint32_t val = fw_peek(p, n, bit, (uint32_t) bits);
        uint32_t lower = (1u << (bits - 1)) - 1u;
        uint32_t thr = (1u << bits) - 1u - ((uint32_t) remaining + 1u);
        int      prob;

        if ((val & lower) < thr)
        {
            val &= lower;
            bit += (uint32_t) bits - 1u;
        }
        else
        {
            if (val > lower)
                val -= thr;
            bit += (uint32_t) bits;
        }
        prob = (int) val - 1;
        remaining -= prob < 0 ? 1 : prob;
        counts[sym++] = (int16_t) prob;
        if (prob == 0)
        {
            uint32_t rep;

            do
            {
                rep = fw_peek(p, n, bit, 2);
                bit += 2;
                for (uint32_t i = 0; i < rep && sym <= max_sym; i++)
                    counts[sym++] = 0;
            } while (rep == 3 && (bit >> 3) <= n);
        }
        if ((bit >> 3) > n)
            return 0;
    }
    if (remaining != 0 || sym > max_sym + 1)
        return 0;
    uint32_t used = (bit + 7u) >> 3;

    if (used > n)
        return 0;
    *nsym_out = sym;
    *log_out = log;
    return used;
}

/* FSE decoding table: cell = symbol | nbits << 8 | base << 16.  Single lane. */
CRYO_DEV void fse_build_table(uint32_t *cell, const int16_t *counts, int nsym, int log,
                              uint16_t *next)
{
    const int size = 1 << log;
    int       high = size - 1, pos = 0;
    const int step = (size >> 1) + (size >> 3) + 3;

    for (int s = 0; s < nsym; s++)
    {
        if (counts[s] == -1)
        {
            cell[high--] = (uint32_t) s;
            next[s] = 1;
        }
        else
            next[s] = (uint16_t) counts[s];
    }
    for (int s = 0; s < nsym; s++)
        for (int i = 0; i < counts[s]; i++)
        {
            cell[pos] = (uint32_t) s;
            do
                pos = (pos + step) & (size - 1);
            while (pos > high);
        }
    for (int i = 0; i < size; i++)
    {
        uint32_t s = cell[i];
        uint32_t nx = next[s]++;
        uint32_t nb = (uint32_t) (log - zs_highbit(nx));

        cell[i] = s | (nb << 8) | ((((nx << nb) - (uint32_t) size) & 0xFFFFu) << 16);
    }
}

/* ---- Huffman (warp 1) ----------------------------------------------------- */

/*
 * Tree description -> decoding table huf[1 << log] (u16: symbol | nbits << 8).
 * Executed by one full warp.  Returns bytes used by the description, 0 on error.
 */
CRYO_DEV uint32_t huf_build_table(const uint8_t *src, uint32_t n, uint16_t *huf, uint8_t *work,
                                  int32_t *log_out, uint32_t lane)
{
    uint8_t  *weights = work + ZW_WEIGHTS;
    uint16_t *symstart = reinterpret_cast<uint16_t *>(work + ZW_SYMSTART);
    uint32_t *wfse = reinterpret_cast<uint32_t *>(work + ZW_WFSE);
    int16_t  *wcounts = reinterpret_cast<int16_t *>(work + ZW_WCOUNTS);
    uint16_t *wnext = reinterpret_cast<uint16_t *>(work + ZW_NEXT + 3 * 128);
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
        return 0;
    int log = zs_highbit(sum) + 1;

    if (log > 11)
        return 0;
    uint32_t left = (1u << log) - sum;

    if (left & (left - 1))
        return 0;
    if (lane == 0)
    {
        weights[nw] = (uint8_t) (zs_highbit(left) + 1);
        /* first cell of every symbol: cells ordered by ascending weight, then symbol */
        uint32_t rank_count[13], rank_start[14];

        for (int r = 0; r < 13; r++)
            rank_count[r] = 0;
        for (uint32_t s = 0; s <= nw; s++)
            rank_count[weights[s]]++;
        rank_start[1] = 0;
        for (int r = 1; r <= log; r++)
            rank_start[r + 1] = rank_start[r] + (rank_count[r] << (r - 1));
        for (uint32_t s = 0; s <= nw; s++)
        {
            uint32_t w = weights[s];

            if (w)
            {
                symstart[s] = (uint16_t) rank_start[w];
                rank_start[w] += 1u << (w - 1);
            }
        }
    }
    nw += 1;
    __syncwarp();
    for (uint32_t s = 0; s < nw; s++)
    {
        uint32_t w = weights[s];

        if (w == 0)
            continue;
        uint32_t len = 1u << (w - 1), st = symstart[s];
        uint16_t ent = (uint16_t) (s | ((uint32_t) (log + 1 - (int) w) << 8));

        for (uint32_t i = lane; i < len; i += 32)
            huf[st + i] = ent;
    }
    __syncwarp();
    *log_out = log;
    return used;
}
