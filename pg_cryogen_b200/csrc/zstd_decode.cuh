/*
 * zstd_decode.cuh -- batched zstd *frame* decompression (RFC 8878), one CTA per
 * cryo block.
 *
 * Replaces ZSTD_decompress as called by the reference at compression.c:116
 * (zstd_decompress, compression.c:111-123): standard frame(s) in, at most `cap`
 * bytes out, non-zero status on any malformed input.  Handles everything
 * ZSTD_compress emits at any level plus what libzstd accepts beyond that
 * (SURVEY.md 8(c), D.1): Raw / RLE / Compressed blocks; Raw, RLE, Huffman
 * (1 or 4 streams, FSE-compressed or direct weights) and treeless literals;
 * predefined / RLE / FSE / repeat sequence tables; repeat offsets; windowed and
 * single-segment frames; concatenated and skippable frames.  The optional
 * content checksum is skipped, not verified (ZSTD_compress never writes one).
 *
 * Work split inside the CTA, per Compressed block:
 *   entropy phase   warp 1: Huffman table + the (up to 4) literal streams, one
 *                   lane per stream, into a global scratch line that stays in L2;
 *                   warps 0/2/3: the three FSE decoding tables (LL / OF / ML)
 *   sequence phase  warp 0 decodes the FSE sequence bitstream (all lanes
 *                   redundantly: the state is warp-uniform, loads broadcast) and
 *                   executes each sequence through cryo_exec.cuh as it appears --
 *                   shared-memory tile for short copies, whole-CTA coalesced
 *                   16-byte stores for long ones (RLE blocks, zero runs).
 */
#pragma once
#include "cryo_exec.cuh"

#define ZSTDD_THREADS 128
#define ZS_LITWIN     4096
#define ZS_IN_BYTES   (12 * 1024)
#define ZS_MAXBLOCK   (1u << 17)
#define ZSTDD_SCRATCH_BYTES (ZS_MAXBLOCK + 256)

#define ZS_SM_EXEC    0
#define ZS_SM_CTL     128
#define ZS_SM_TILE    512
#define ZS_SM_PAT     (ZS_SM_TILE + EX_TILE)
#define ZS_SM_HUF     (ZS_SM_PAT + EX_PAT_BYTES)
#define ZS_SM_LL      (ZS_SM_HUF + 4096)
#define ZS_SM_OF      (ZS_SM_LL + 2048)
#define ZS_SM_ML      (ZS_SM_OF + 1024)
#define ZS_SM_WORK    (ZS_SM_ML + 2048)
#define ZS_SM_LITWIN  (ZS_SM_WORK + 2048)
#define ZS_SM_IN      (ZS_SM_LITWIN + ZS_LITWIN)
#define ZSTDD_SMEM    (ZS_SM_IN + ZS_IN_BYTES)

/* work-area offsets (bytes from ZS_SM_WORK) */
#define ZW_WEIGHTS    0       /* u8[256]  Huffman weights */
#define ZW_SYMSTART   256     /* u16[256] first table cell of every symbol */
#define ZW_WFSE       768     /* u32[64]  FSE table of the Huffman weights */
#define ZW_WCOUNTS    1024    /* i16[16] */
#define ZW_COUNTS     1152    /* i16[64] x 3: LL, OF, ML normalised counts */
#define ZW_NEXT       1536    /* u16[64] x 4: per-builder scratch */

#define ZC_ENTROPY    16

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

struct ZstdCtl
{
    /* literals section (Huffman kinds only), set by the master per block */
    const uint8_t *lit_src;     /* tree description (type 2) or streams (type 3) */
    uint8_t    *lit_dst;        /* global scratch, 16-byte aligned */
    uint32_t    lit_csize;
    uint32_t    lit_regen;
    int32_t     lit_type;       /* 2 compressed, 3 treeless, anything else: nothing to do */
    int32_t     lit_streams;
    /* sequence table descriptions */
    const uint8_t *seq_src;
    uint32_t    seq_len;
    int32_t     modes;          /* modes byte, or -1 when the block has no sequences */
    uint32_t    seq_used;       /* out: bytes taken by the table descriptions */
    /* tables that persist across blocks of a frame */
    int32_t     huf_log, ll_log, of_log, ml_log;    /* -1 = not yet defined */
    int32_t     cnt_log[3], cnt_nsym[3];
    int32_t     err;
};

CRYO_DEV int zs_highbit(uint32_t v) { return 31 - __clz((int) v); }

CRYO_DEV void zs_named_barrier(int id, int count)
{
#ifdef CRYO_EMU
    emu_named_barrier(id, count);
#else
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
#endif
}

CRYO_DEV void zs_set_err(ZstdCtl *ctl, int e)
{
    if (ctl->err == ST_OK)
        ctl->err = e;
}

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

/* one Huffman stream, one lane: `count` symbols to dst; returns false on corruption */
CRYO_DEV bool huf_decode_stream(const uint16_t *huf, int log, const uint8_t *src, uint32_t n,
                                uint8_t *dst, uint32_t count)
{
    BitsBack bb;

    if (!bb_init(bb, src, n))
        return false;
    const uint32_t sh = 64u - (uint32_t) log;

    for (uint32_t i = 0; i < count; i++)
    {
        bb_refill(bb);
        uint32_t ent = huf[(uint32_t) (bb.acc >> sh)];
        uint32_t nb = ent >> 8;

        dst[i] = (uint8_t) ent;
        bb.acc <<= nb;
        bb.avail -= (int32_t) nb;
        bb.remaining -= (int32_t) nb;
    }
    return bb.remaining == 0;
}

CRYO_DEV void zs_literals_warp(ZstdCtl *ctl, uint16_t *huf, uint8_t *work, uint32_t lane)
{
    const uint8_t *p = ctl->lit_src;
    uint32_t left = ctl->lit_csize, regen = ctl->lit_regen;

    if (ctl->lit_type == 2)
    {
        int32_t  log = 0;
        uint32_t used = huf_build_table(p, left, huf, work, &log, lane);

        if (used == 0)
        {
            if (lane == 0)
            {
                zs_set_err(ctl, ST_FORMAT);
                ctl->huf_log = -1;
            }
            return;
        }
        if (lane == 0)
            ctl->huf_log = log;
        p += used;
        left -= used;
    }
    __syncwarp();
    int log = ctl->huf_log;

    if (log < 1)
    {
        if (lane == 0)
            zs_set_err(ctl, ST_FORMAT);
        return;
    }
    bool ok = true;

    if (ctl->lit_streams == 1)
    {
        if (lane == 0)
            ok = huf_decode_stream(huf, log, p, left, ctl->lit_dst, regen);
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

                ok = huf_decode_stream(huf, log, p + 6 + so, sn, ctl->lit_dst + lane * seg, cnt);
            }
        }
    }
    if (__any_sync(CRYO_FULL, !ok) && lane == 0)
        zs_set_err(ctl, ST_FORMAT);
}

/* ---- sequence tables (warps 0, 2, 3) -------------------------------------- */

/* lane 0 of warp 2: walk the three table descriptions in stream order */
CRYO_DEV void zs_read_seq_descriptions(ZstdCtl *ctl, uint8_t *work)
{
    const uint8_t *p = ctl->seq_src;
    uint32_t left = ctl->seq_len, used_total = 0;
    const int max_log[3] = {9, 8, 9}, max_sym[3] = {35, 31, 52};
    int16_t  *counts = reinterpret_cast<int16_t *>(work + ZW_COUNTS);

    for (int t = 0; t < 3; t++)
    {
        int mode = (ctl->modes >> (6 - 2 * t)) & 3;

        if (mode == 1)
        {
            if (left < 1 || p[0] > max_sym[t])
            {
                zs_set_err(ctl, ST_FORMAT);
                return;
            }
            counts[64 * t] = (int16_t) p[0];
            p += 1;
            left -= 1;
            used_total += 1;
        }
        else if (mode == 2)
        {
            uint32_t u = fse_read_counts(p, left, max_log[t], max_sym[t], counts + 64 * t,
                                         &ctl->cnt_nsym[t], &ctl->cnt_log[t]);

            if (u == 0)
            {
                zs_set_err(ctl, ST_FORMAT);
                return;
            }
            p += u;
            left -= u;
            used_total += u;
        }
    }
    ctl->seq_used = used_total;
}

/* lane 0 of one warp: build table t (0 LL, 1 OF, 2 ML) according to its mode */
CRYO_DEV void zs_build_seq_table(ZstdCtl *ctl, int t, uint32_t *cell, uint8_t *work)
{
    int      mode = (ctl->modes >> (6 - 2 * t)) & 3;
    int16_t *counts = reinterpret_cast<int16_t *>(work + ZW_COUNTS) + 64 * t;
    uint16_t *next = reinterpret_cast<uint16_t *>(work + ZW_NEXT) + 64 * t;
    int32_t *logp = t == 0 ? &ctl->ll_log : t == 1 ? &ctl->of_log : &ctl->ml_log;

    if (ctl->err != ST_OK)
        return;
    switch (mode)
    {
        case 0:
        {
            int     n = t == 0 ? 36 : t == 1 ? 29 : 53;
            const int16_t *def = t == 0 ? ZS_LL_DEFAULT : t == 1 ? ZS_OF_DEFAULT : ZS_ML_DEFAULT;

            fse_build_table(cell, def, n, t == 1 ? 5 : 6, next);
            *logp = t == 1 ? 5 : 6;
            break;
        }
        case 1:
            cell[0] = (uint32_t) (uint16_t) counts[0];     /* nbits 0, base 0 */
            *logp = 0;
            break;
        case 2:
            fse_build_table(cell, counts, ctl->cnt_nsym[t], ctl->cnt_log[t], next);
            *logp = ctl->cnt_log[t];
            break;
        default:
            if (*logp < 0)
                zs_set_err(ctl, ST_FORMAT);
            break;
    }
}

/* the entropy phase of one Compressed block; every thread of the CTA calls it */
CRYO_DEV void zs_entropy_phase(ZstdCtl *ctl, uint8_t *smem, uint32_t tid)
{
    const uint32_t warp = tid >> 5, lane = tid & 31;
    uint8_t *work = smem + ZS_SM_WORK;

    if (warp == 1)
    {
        if (ctl->lit_type >= 2)
            zs_literals_warp(ctl, reinterpret_cast<uint16_t *>(smem + ZS_SM_HUF), work, lane);
        return;
    }
    if (ctl->modes < 0)
        return;
    if (warp == 2 && lane == 0)
        zs_read_seq_descriptions(ctl, work);
    zs_named_barrier(1, 96);
    if (lane == 0)
    {
        if (warp == 0)
            zs_build_seq_table(ctl, 0, reinterpret_cast<uint32_t *>(smem + ZS_SM_LL), work);
        else if (warp == 2)
            zs_build_seq_table(ctl, 1, reinterpret_cast<uint32_t *>(smem + ZS_SM_OF), work);
        else
            zs_build_seq_table(ctl, 2, reinterpret_cast<uint32_t *>(smem + ZS_SM_ML), work);
    }
}

CRYO_DEV void zs_handle(uint8_t *out, uint8_t *smem, const ExecShared *sh, ZstdCtl *ctl,
                        uint32_t tid, uint32_t nthr)
{
    if (sh->op == ZC_ENTROPY)
        zs_entropy_phase(ctl, smem, tid);
    else
        exec_handle(out, smem + ZS_SM_TILE, smem + ZS_SM_PAT, sh, tid, nthr);
}

/* master-side issue that understands ZC_ENTROPY */
CRYO_DEV void zs_issue_entropy(Exec &e, ZstdCtl *ctl, uint8_t *smem, uint32_t tid, uint32_t nthr)
{
    if (tid == 0)
    {
        e.sh->op = ZC_ENTROPY;
        e.sh->tbase = e.tbase;
        e.sh->flushed = e.flushed;
        e.sh->pos = e.pos;
    }
    __syncthreads();
    zs_entropy_phase(ctl, smem, tid);
    __syncthreads();
}

/* ---- literal source for the sequence executor ------------------------------ */

struct ZsLits
{
    int         kind;           /* 0 bytes at `base`, 1 RLE */
    bool        direct;         /* base is shared memory: no staging window needed */
    uint8_t     rle;
    const uint8_t *base;        /* first literal byte */
    uint32_t    n;              /* literal count of the block */
    uint32_t    pos;            /* literals consumed */
    /* staging window over global sources */
    const uint8_t *abase;       /* 16-byte aligned address at or before base */
    uint32_t    delta;          /* base - abase */
    uint32_t    wbase;          /* in abase coordinates, multiple of 16 */
    bool        wvalid;
    uint8_t    *win;
};

CRYO_DEV void zs_lits_emit(Exec &e, ZsLits &L, uint32_t n, uint32_t tid, uint32_t nthr)
{
    if (n == 0)
        return;
    if (L.kind == 1)
    {
        if (n >= EX_BULK)
            exec_fill_byte(e, L.rle, n, tid, nthr);
        else
            exec_fill_small(e, L.rle, n, tid, nthr);
    }
    else if (n >= EX_BULK)
    {
        exec_issue(e, EXC_COPY, n, 0, L.base + L.pos, tid, nthr);
        exec_after_bulk(e, n, tid);
    }
    else if (L.direct)
        exec_literals_small(e, L.base + L.pos, n, tid, nthr);
    else
    {
        uint32_t ip = L.delta + L.pos;

        if (!L.wvalid || ip < L.wbase || ip + n > L.wbase + ZS_LITWIN)
        {
            uint32_t lim = (L.delta + L.n + 15u) & ~15u;

            __syncwarp();
            L.wbase = align_down16(ip);
            L.wvalid = true;
#pragma unroll 4
            for (uint32_t v = tid; v < ZS_LITWIN / 16; v += 32)
            {
                uint32_t a = L.wbase + 16 * v;

                if (a < lim)
                    st16(L.win + 16 * v, ld16(L.abase + a));
            }
            __syncwarp();
        }
        exec_literals_small(e, L.win + (ip - L.wbase), n, tid, nthr);
    }
    L.pos += n;
}

/* ---- the frame decoder ------------------------------------------------------ */

/*
 * Decode the zstd frame(s) at src[0, csize) into out[0, cap).  Called by every
 * thread of the CTA.  scratch: ZSTDD_SCRATCH_BYTES of global memory private to
 * this CTA, 16-byte aligned.
 */
CRYO_DEV void zstd_decode_frame(const uint8_t *src, uint32_t csize, uint8_t *out, uint32_t cap,
                                uint32_t *out_size, int32_t *status, uint8_t *scratch)
{
    uint8_t    *smem = CRYO_SMEM_BASE();
    ExecShared *sh = reinterpret_cast<ExecShared *>(smem + ZS_SM_EXEC);
    ZstdCtl    *ctl = reinterpret_cast<ZstdCtl *>(smem + ZS_SM_CTL);
    uint8_t    *tile = smem + ZS_SM_TILE;
    uint8_t    *pat = smem + ZS_SM_PAT;
    const uint32_t tid = threadIdx.x, nthr = ZSTDD_THREADS;

    /* small frames (every sparse cryo block) are parsed out of shared memory */
    const uint8_t *in = src;

    if (csize <= ZS_IN_BYTES - 32)
    {
        uint32_t d = (uint32_t) ((uintptr_t) src & 15u);

        team_copy(smem + ZS_SM_IN + d, src, csize, tid, nthr);
        in = smem + ZS_SM_IN + d;
    }
    if (tid == 0)
    {
        ctl->err = ST_OK;
        ctl->huf_log = ctl->ll_log = ctl->of_log = ctl->ml_log = -1;
    }
    __syncthreads();

    if (tid >= 32)
    {
        for (;;)
        {
            __syncthreads();
            if (sh->op == EXC_EXIT)
                break;
            zs_handle(out, smem, sh, ctl, tid, nthr);
            __syncthreads();
        }
        return;
    }

    /* ---- master warp ---- */
    Exec     e;
    int      err = ST_OK;
    uint32_t ip = 0;
    uint32_t *ll_tab = reinterpret_cast<uint32_t *>(smem + ZS_SM_LL);
    uint32_t *of_tab = reinterpret_cast<uint32_t *>(smem + ZS_SM_OF);
    uint32_t *ml_tab = reinterpret_cast<uint32_t *>(smem + ZS_SM_ML);

    exec_init(e, out, cap, tile, pat, sh);
    /* csize == 0: ZSTD_decompress returns 0 (no frame, no error) and the reference
     * accepts it (compression.c:116-118); so do we */

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
        uint64_t window = 0, fcs = 0;

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
            uint32_t b = in[ip++], wl = 10 + (b >> 3);

            if (wl > 27)
            {
                err = ST_FORMAT;
                break;
            }
            window = (1ull << wl) + ((1ull << wl) / 8) * (b & 7u);
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
        if (single)
            window = fcs;
        /* RFC 8878 says min(Window_Size, 128 KiB); libzstd 1.5.5's ZSTD_decompress (the
         * reference's call, compression.c:116) only enforces the constant -- follow it */
        const uint32_t block_max = ZS_MAXBLOCK;
        (void) window;
        const uint32_t frame_start = e.pos;
        uint32_t rep0 = 1, rep1 = 4, rep2 = 8;

        if (tid == 0)
            ctl->huf_log = ctl->ll_log = ctl->of_log = ctl->ml_log = -1;
        __syncwarp();

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
                if (bsize > cap - e.pos)
                {
                    err = ST_OUTPUT;
                    break;
                }
                if (bsize)
                    exec_literals(e, in + ip, nullptr, bsize, tid, nthr);
                ip += bsize;
            }
            else if (type == 1)
            {
                if (ip + 1 > csize)
                {
                    err = ST_INPUT;
                    break;
                }
                if (bsize > cap - e.pos)
                {
                    err = ST_OUTPUT;
                    break;
                }
                if (bsize >= 64)
                    exec_fill_byte(e, in[ip], bsize, tid, nthr);
                else if (bsize)
                    exec_fill_small(e, in[ip], bsize, tid, nthr);
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
                const uint32_t block_start = e.pos;
                uint32_t lt = bp[0] & 3u, sf = (bp[0] >> 2) & 3u;
                uint32_t lhdr, regen, lcsize = 0, streams = 1;
                ZsLits   L;

                L.pos = 0;
                L.wvalid = false;
                L.win = smem + ZS_SM_LITWIN;
                L.direct = (in != src);
                L.kind = 0;
                L.rle = 0;
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
                    L.base = bp + lhdr;
                    if (lt == 1)
                    {
                        L.kind = 1;
                        L.rle = bp[lhdr];
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
                    L.base = scratch;
                    L.direct = false;
                }
                L.n = regen;
                L.abase = L.base - ((uintptr_t) L.base & 15u);
                L.delta = (uint32_t) ((uintptr_t) L.base & 15u);

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
                int modes = -1;

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
                /* entropy phase: Huffman literals on warp 1, FSE tables on warps 0/2/3 */
                if (lt >= 2 || nseq)
                {
                    if (tid == 0)
                    {
                        ctl->lit_src = bp + lhdr;
                        ctl->lit_dst = scratch;
                        ctl->lit_csize = lcsize;
                        ctl->lit_regen = regen;
                        ctl->lit_type = (int32_t) lt;
                        ctl->lit_streams = (int32_t) streams;
                        ctl->seq_src = bp + sp;
                        ctl->seq_len = bsize - sp;
                        ctl->modes = modes;
                        ctl->seq_used = 0;
                    }
                    zs_issue_entropy(e, ctl, smem, tid, nthr);
                    err = ctl->err;
                    if (err != ST_OK)
                        break;
                }
                if (nseq)
                {
                    BitsBack bb;
                    uint32_t tp = sp + ctl->seq_used;
                    int      ll_log = ctl->ll_log, of_log = ctl->of_log, ml_log = ctl->ml_log;

                    if (tp > bsize || !bb_init(bb, bp + tp, bsize - tp))
                    {
                        err = ST_FORMAT;
                        break;
                    }
                    bb_refill(bb);
                    uint32_t sl = bb_read(bb, (uint32_t) ll_log);
                    uint32_t so = bb_read(bb, (uint32_t) of_log);

                    bb_refill(bb);
                    uint32_t sm = bb_read(bb, (uint32_t) ml_log);

                    for (uint32_t i = 0; i < nseq; i++)
                    {
                        uint32_t cl = ll_tab[sl], co = of_tab[so], cm = ml_tab[sm];
                        uint32_t lc = cl & 0xFFu, oc = co & 0xFFu, mc = cm & 0xFFu;

                        if (lc > 35 || mc > 52 || oc > 31)
                        {
                            err = ST_FORMAT;
                            break;
                        }
                        bb_refill(bb);
                        uint32_t ov = (1u << oc) + bb_read(bb, oc);

                        bb_refill(bb);
                        uint32_t ml = ZS_ML_BASE[mc] + bb_read(bb, ZS_ML_BITS[mc]);
                        uint32_t ll = ZS_LL_BASE[lc] + bb_read(bb, ZS_LL_BITS[lc]);
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
                            uint32_t idx = ov - 1 + (ll == 0 ? 1u : 0u);

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
                        if (i + 1 < nseq)
                        {
                            bb_refill(bb);
                            sl = (cl >> 16) + bb_read(bb, (cl >> 8) & 0xFFu);
                            sm = (cm >> 16) + bb_read(bb, (cm >> 8) & 0xFFu);
                            so = (co >> 16) + bb_read(bb, (co >> 8) & 0xFFu);
                        }
                        if (bb.remaining < 0)
                        {
                            err = ST_INPUT;
                            break;
                        }
                        if (ll > L.n - L.pos)
                        {
                            err = ST_FORMAT;
                            break;
                        }
                        if ((uint64_t) e.pos + ll + ml > cap)
                        {
                            err = ST_OUTPUT;
                            break;
                        }
                        if (e.pos + ll + ml - block_start > block_max)
                        {
                            err = ST_FORMAT;
                            break;
                        }
                        zs_lits_emit(e, L, ll, tid, nthr);
                        if (off == 0 || off > e.pos - frame_start)
                        {
                            err = ST_OFFSET;
                            break;
                        }
                        exec_match(e, off, ml, tid, nthr);
                    }
                    if (err != ST_OK)
                        break;
                    if (bb.remaining != 0)
                    {
                        err = ST_INPUT;
                        break;
                    }
                }
                /* literals left after the last sequence */
                uint32_t rest = L.n - L.pos;

                if (rest > cap - e.pos)
                {
                    err = ST_OUTPUT;
                    break;
                }
                if (e.pos + rest - block_start > block_max)
                {
                    err = ST_FORMAT;
                    break;
                }
                zs_lits_emit(e, L, rest, tid, nthr);
                ip += bsize;
            }
            if (last)
                break;
        }
        if (err != ST_OK)
            break;
        if (fcs_bytes && (uint64_t) (e.pos - frame_start) != fcs)
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
    exec_finish(e, tid, nthr);
    if (tid == 0)
    {
        *out_size = err == ST_OK ? e.pos : 0u;
        *status = err;
    }
}
