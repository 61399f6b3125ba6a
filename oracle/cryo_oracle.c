/*
 * oracle/cryo_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of the LZ4 block format and the zstd frame format
 * (RFC 8878) as the reference consumes them through LZ4_decompress_safe
 * (compression.c:84) and ZSTD_decompress (compression.c:116).  Written from the
 * published format descriptions (summarised in SURVEY.md appendix B); it calls
 * no library.  See cryo_oracle.h for how it is pinned.
 */
#include "cryo_oracle.h"

#include <string.h>

#define ERR(e) return (e)

/* ------------------------------------------------------------------ LZ4 -- */

/*
 * LZ4 block: sequences of
 *   token (hi nibble = literal length, lo nibble = match length - 4)
 *   [literal length extension bytes while 255]  literals
 *   offset (u16 LE)  [match length extension bytes while 255]
 * The last sequence stops after its literals.
 *
 * Acceptance rules follow LZ4_decompress_safe (liblz4 1.9.4) as probed in
 * input must be consumed exactly; a literal run that ends
 * within 12 bytes of the output capacity or within 8 bytes of the input end
 * must be the final one; a match must end at least 5 bytes before the output
 * capacity; offsets before the start of output are rejected.  Offset 0 is
 * format-invalid; liblz4 does not reject it, this port does (parity is only
 * required on valid streams).
 */
long
cryo_oracle_lz4_decode(const uint8_t *src, size_t src_size, uint8_t *dst, size_t dst_cap,
                       cryo_oracle_stats *st, cryo_oracle_seq_cb cb, void *cb_ctx)
{
    size_t ip = 0, op = 0;

    if (st)
        memset(st, 0, sizeof(*st));
    if (src_size == 0)
        ERR(CRYO_ORACLE_ERR_INPUT);
    for (;;)
    {
        uint32_t token, ll, ml, off;

        if (ip >= src_size)
            ERR(CRYO_ORACLE_ERR_INPUT);
        token = src[ip++];
        ll = token >> 4;
        if (ll == 15)
        {
            uint32_t b;

            do
            {
                if (ip >= src_size)
                    ERR(CRYO_ORACLE_ERR_INPUT);
                b = src[ip++];
                ll += b;
            } while (b == 255);
        }
        if (ip + ll > src_size)
            ERR(CRYO_ORACLE_ERR_INPUT);
        if (op + ll > dst_cap)
            ERR(CRYO_ORACLE_ERR_OUTPUT);
        /* a run this close to either end must be the last one */
        if (op + ll + 12 > dst_cap || ip + ll + 8 > src_size)
        {
            if (ip + ll != src_size)
                ERR(ip + ll + 8 > src_size && op + ll + 12 <= dst_cap ? CRYO_ORACLE_ERR_INPUT
                                                                      : CRYO_ORACLE_ERR_OUTPUT);
        }
        memcpy(dst + op, src + ip, ll);
        ip += ll;
        op += ll;
        if (st)
        {
            st->literal_bytes += ll;
            if (ll > st->max_literal_run)
                st->max_literal_run = ll;
        }
        if (ip == src_size)
        {
            if (cb)
                cb(cb_ctx, ll, 0, 0);
            break;
        }
        off = src[ip] | ((uint32_t) src[ip + 1] << 8);   /* ip + 2 <= src_size: checked above (8 spare) */
        ip += 2;
        ml = token & 15;
        if (ml == 15)
        {
            uint32_t b;

            do
            {
                if (ip >= src_size)
                    ERR(CRYO_ORACLE_ERR_INPUT);
                b = src[ip++];
                ml += b;
            } while (b == 255);
        }
        ml += 4;
        if (off == 0 || off > op)
            ERR(CRYO_ORACLE_ERR_OFFSET);
        if (op + ml + 5 > dst_cap)
            ERR(CRYO_ORACLE_ERR_OUTPUT);
        for (uint32_t i = 0; i < ml; i++)        /* byte-serial: overlap defines a periodic fill */
            dst[op + i] = dst[op + i - off];
        op += ml;
        if (st)
        {
            st->sequences++;
            st->match_bytes += ml;
            if (ml > st->longest_match)
                st->longest_match = ml;
            if (off < ml)
                st->overlap_match_bytes += ml;
            if (off > st->max_offset)
                st->max_offset = off;
        }
        if (cb)
            cb(cb_ctx, ll, ml, off);
    }
    return (long) op;
}

/* ---------------------------------------------------------------- XXH64 -- */

#define P1 0x9E3779B185EBCA87ULL
#define P2 0xC2B2AE3D27D4EB4FULL
#define P3 0x165667B19E3779F9ULL
#define P4 0x85EBCA77C2B2AE63ULL
#define P5 0x27D4EB2F165667C5ULL

static uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static uint64_t rd64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
static uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint64_t xxround(uint64_t acc, uint64_t in) { return rotl64(acc + in * P2, 31) * P1; }
static uint64_t xxmerge(uint64_t h, uint64_t v) { return (h ^ xxround(0, v)) * P1 + P4; }

uint64_t
cryo_oracle_xxh64(const uint8_t *p, size_t n, uint64_t seed)
{
    const uint8_t *end = p + n;
    uint64_t h;

    if (n >= 32)
    {
        uint64_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;

        do
        {
            v1 = xxround(v1, rd64(p));
            v2 = xxround(v2, rd64(p + 8));
            v3 = xxround(v3, rd64(p + 16));
            v4 = xxround(v4, rd64(p + 24));
            p += 32;
        } while (p + 32 <= end);
        h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
        h = xxmerge(h, v1);
        h = xxmerge(h, v2);
        h = xxmerge(h, v3);
        h = xxmerge(h, v4);
    }
    else
        h = seed + P5;
    h += n;
    while (p + 8 <= end)
    {
        h ^= xxround(0, rd64(p));
        h = rotl64(h, 27) * P1 + P4;
        p += 8;
    }
    if (p + 4 <= end)
    {
        h ^= (uint64_t) rd32(p) * P1;
        h = rotl64(h, 23) * P2 + P3;
        p += 4;
    }
    while (p < end)
    {
        h ^= (*p++) * P5;
        h = rotl64(h, 11) * P1;
    }
    h ^= h >> 33;
    h *= P2;
    h ^= h >> 29;
    h *= P3;
    h ^= h >> 32;
    return h;
}

/* ----------------------------------------------------------------- zstd -- */

static int
highbit(uint32_t v)             /* index of the highest set bit; v != 0 */
{
    return 31 - __builtin_clz(v);
}

/* forward (LSB-first) bit reader, used for FSE table descriptions */
typedef struct
{
    const uint8_t *p;
    size_t      n;
    size_t      bit;            /* next bit index */
} fbits;

static uint32_t
fpeek(const fbits *b, int k)
{
    uint32_t v = 0;

    for (int i = 0; i < k; i++)
    {
        size_t pos = b->bit + i;

        if ((pos >> 3) < b->n)
            v |= (uint32_t) ((b->p[pos >> 3] >> (pos & 7)) & 1) << i;
    }
    return v;
}

/*
 * backward bit reader: the stream is a little-endian integer; the highest set
 * bit of the last byte is the end marker; values are taken from the top down.
 * Reading below bit 0 yields zero bits (and is remembered as overflow).
 */
typedef struct
{
    const uint8_t *p;
    long        bit;            /* number of unread bits; may go negative */
} bbits;

static int
binit(bbits *b, const uint8_t *p, size_t n)
{
    if (n == 0 || p[n - 1] == 0)
        return -1;
    b->p = p;
    b->bit = (long) (n - 1) * 8 + highbit(p[n - 1]);
    return 0;
}

static uint32_t
bread(bbits *b, int k)          /* k <= 32 */
{
    uint64_t v = 0;

    b->bit -= k;
    for (int i = 0; i < k; i++)
    {
        long pos = b->bit + i;

        if (pos >= 0)
            v |= (uint64_t) ((b->p[pos >> 3] >> (pos & 7)) & 1) << i;
    }
    return (uint32_t) v;
}

/* --- FSE --- */

typedef struct
{
    uint8_t     symbol;
    uint8_t     nbits;
    uint16_t    base;
} fse_cell;

typedef struct
{
    int         log;            /* accuracy log; table has 1 << log cells */
    fse_cell    cell[1 << 9];
} fse_table;

/* read a normalised-count description (RFC 8878 4.1.1); returns bytes used or <0 */
static long
fse_read_counts(const uint8_t *src, size_t n, int max_log, int max_sym, int16_t *counts,
                int *nsym, int *log_out)
{
    fbits fb = {src, n, 0};
    int log, remaining, sym = 0;

    if (n == 0)
        return -1;
    log = (int) fpeek(&fb, 4) + 5;
    fb.bit += 4;
    if (log > max_log)
        return -1;
    remaining = 1 << log;
    while (remaining > 0 && sym <= max_sym)
    {
        int bits = highbit((uint32_t) remaining + 1) + 1;
        uint32_t val = fpeek(&fb, bits);
        uint32_t lower = (1u << (bits - 1)) - 1;
        uint32_t thr = (1u << bits) - 1 - ((uint32_t) remaining + 1);
        int prob;

        if ((val & lower) < thr)
        {
            val &= lower;
            fb.bit += bits - 1;
        }
        else
        {
            if (val > lower)
                val -= thr;
            fb.bit += bits;
        }
        prob = (int) val - 1;
        remaining -= prob < 0 ? 1 : prob;
        counts[sym++] = (int16_t) prob;
        if (prob == 0)
        {
            uint32_t rep;

            do
            {
                rep = fpeek(&fb, 2);
                fb.bit += 2;
                for (uint32_t i = 0; i < rep && sym <= max_sym; i++)
                    counts[sym++] = 0;
            } while (rep == 3);
        }
    }
    if (remaining != 0 || sym > max_sym + 1)
        return -1;
    if ((fb.bit + 7) / 8 > n)
        return -1;
    *nsym = sym;
    *log_out = log;
    return (long) ((fb.bit + 7) / 8);
}

/* build the decoding table from normalised counts (RFC 8878 4.1.1) */
static void
fse_build(fse_table *t, const int16_t *counts, int nsym, int log)
{
    int size = 1 << log, high = size - 1, pos = 0;
    int step = (size >> 1) + (size >> 3) + 3;
    uint16_t next[256];

    t->log = log;
    for (int s = 0; s < nsym; s++)
        if (counts[s] == -1)
        {
            t->cell[high--].symbol = (uint8_t) s;
            next[s] = 1;
        }
        else
            next[s] = (uint16_t) counts[s];
    for (int s = 0; s < nsym; s++)
        for (int i = 0; i < counts[s]; i++)
        {
            t->cell[pos].symbol = (uint8_t) s;
            do
                pos = (pos + step) & (size - 1);
            while (pos > high);
        }
    for (int i = 0; i < size; i++)
    {
        int s = t->cell[i].symbol;
        uint16_t nx = next[s]++;
        int nb = log - highbit(nx);

        t->cell[i].nbits = (uint8_t) nb;
        t->cell[i].base = (uint16_t) ((nx << nb) - size);
    }
}

static void
fse_build_rle(fse_table *t, uint8_t sym)
{
    t->log = 0;
    t->cell[0].symbol = sym;
    t->cell[0].nbits = 0;
    t->cell[0].base = 0;
}

/* predefined distributions, RFC 8878 3.1.1.3.2.2 */
static const int16_t LL_DEFAULT[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2,
    2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
static const int16_t ML_DEFAULT[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1,
    1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1,
    -1, -1, -1, -1, -1};
static const int16_t OF_DEFAULT[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1,
    1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};

static const uint32_t LL_BASE[36] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16,
    18, 20, 22, 24, 28, 32, 40, 48, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768,
    65536};
static const uint8_t LL_BITS[36] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1,
    1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
static const uint32_t ML_BASE[53] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18,
    19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 37, 39, 41, 43, 47, 51,
    59, 67, 83, 99, 131, 259, 515, 1027, 2051, 4099, 8195, 16387, 32771, 65539};
static const uint8_t ML_BITS[53] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11,
    12, 13, 14, 15, 16};

/* --- Huffman --- */

typedef struct
{
    int         log;            /* max code length, <= 11 */
    uint8_t     symbol[1 << 11];
    uint8_t     nbits[1 << 11];
} huf_table;

/* read a Huffman tree description (RFC 8878 4.2.1); returns bytes used or <0 */
static long
huf_read_table(const uint8_t *src, size_t n, huf_table *t, cryo_oracle_stats *st)
{
    uint8_t  w[256];
    int      nw = 0;
    long     used;
    uint32_t sum = 0, left;
    int      log, rank_count[13] = {0};
    uint32_t rank_start[14];

    if (n == 0)
        return -1;
    if (src[0] >= 128)
    {
        /* direct: 4 bits per weight, high nibble first */
        nw = src[0] - 127;
        used = 1 + (nw + 1) / 2;
        if ((size_t) used > n)
            return -1;
        for (int i = 0; i < nw; i++)
            w[i] = (i & 1) ? (src[1 + i / 2] & 15) : (src[1 + i / 2] >> 4);
        if (st)
            st->huf_direct_weights++;
    }
    else
    {
        /* FSE-compressed weights, two interleaved states */
        int16_t   counts[16];
        int       nsym, flog;
        long      hdr;
        fse_table ft;
        bbits     bb;
        uint32_t  s1, s2;
        size_t    clen = src[0];

        used = 1 + (long) clen;
        if ((size_t) used > n || clen == 0)
            return -1;
        hdr = fse_read_counts(src + 1, clen, 6, 12, counts, &nsym, &flog);
        if (hdr < 0 || (size_t) hdr >= clen)
            return -1;
        fse_build(&ft, counts, nsym, flog);
        if (binit(&bb, src + 1 + hdr, clen - hdr) < 0)
            return -1;
        s1 = bread(&bb, flog);
        s2 = bread(&bb, flog);
        for (;;)
        {
            if (nw > 253)
                return -1;
            w[nw++] = ft.cell[s1].symbol;
            s1 = ft.cell[s1].base + bread(&bb, ft.cell[s1].nbits);
            if (bb.bit < 0)
            {
                w[nw++] = ft.cell[s2].symbol;
                break;
            }
            w[nw++] = ft.cell[s2].symbol;
            s2 = ft.cell[s2].base + bread(&bb, ft.cell[s2].nbits);
            if (bb.bit < 0)
            {
                w[nw++] = ft.cell[s1].symbol;
                break;
            }
        }
        if (st)
            st->huf_fse_weights++;
    }
    for (int i = 0; i < nw; i++)
    {
        if (w[i] > 11)
            return -1;
        if (w[i])
            sum += 1u << (w[i] - 1);
    }
    if (sum == 0)
        return -1;
    log = highbit(sum) + 1;
    if (log > 11)
        return -1;
    left = (1u << log) - sum;
    if (left & (left - 1))      /* the implied last weight must complete a power of two */
        return -1;
    w[nw++] = (uint8_t) (highbit(left) + 1);
    /* cells by ascending weight (longest codes first), symbols in natural order */
    for (int i = 0; i < nw; i++)
        rank_count[w[i]]++;
    rank_start[1] = 0;
    for (int r = 1; r <= log; r++)
        rank_start[r + 1] = rank_start[r] + ((uint32_t) rank_count[r] << (r - 1));
    for (int s = 0; s < nw; s++)
        if (w[s])
        {
            uint32_t len = 1u << (w[s] - 1);

            for (uint32_t i = 0; i < len; i++)
            {
                t->symbol[rank_start[w[s]] + i] = (uint8_t) s;
                t->nbits[rank_start[w[s]] + i] = (uint8_t) (log + 1 - w[s]);
            }
            rank_start[w[s]] += len;
        }
    t->log = log;
    return used;
}

static int
huf_decode_stream(const huf_table *t, const uint8_t *src, size_t n, uint8_t *dst, size_t count)
{
    bbits    bb;
    uint32_t window;            /* next `log` bits, refreshed incrementally */

    if (binit(&bb, src, n) < 0)
        return -1;
    window = bread(&bb, t->log);
    for (size_t i = 0; i < count; i++)
    {
        int nb = t->nbits[window];

        dst[i] = t->symbol[window];
        window = ((window << nb) & ((1u << t->log) - 1)) | bread(&bb, nb);
    }
    /* all bits consumed exactly: the `log` bits of look-ahead are all below bit 0 */
    if (bb.bit != -(long) t->log)
        return -1;
    return 0;
}

/* --- frame state --- */

typedef struct
{
    huf_table   huf;
    int         huf_valid;
    fse_table   ll, of, ml;
    int         ll_valid, of_valid, ml_valid;
    uint32_t    rep[3];
    uint8_t     lit[1 << 17];
} zstd_ctx;

static long
decode_literals(zstd_ctx *z, const uint8_t *src, size_t n, size_t *regen_out,
                cryo_oracle_stats *st)
{
    int     type, sf;
    size_t  hdr, regen, csize = 0;
    int     streams = 1;

    if (n < 1)
        return -1;
    type = src[0] & 3;
    sf = (src[0] >> 2) & 3;
    if (type < 2)
    {
        if (sf == 0 || sf == 2)
        {
            hdr = 1;
            regen = src[0] >> 3;
        }
        else if (sf == 1)
        {
            if (n < 2)
                return -1;
            hdr = 2;
            regen = (src[0] >> 4) | ((size_t) src[1] << 4);
        }
        else
        {
            if (n < 3)
                return -1;
            hdr = 3;
            regen = (src[0] >> 4) | ((size_t) src[1] << 4) | ((size_t) src[2] << 12);
        }
        if (regen > sizeof(z->lit))
            return -1;
        if (type == 0)
        {
            if (hdr + regen > n)
                return -1;
            memcpy(z->lit, src + hdr, regen);
            if (st)
                st->lit_raw++;
            *regen_out = regen;
            return (long) (hdr + regen);
        }
        if (hdr + 1 > n)
            return -1;
        memset(z->lit, src[hdr], regen);
        if (st)
            st->lit_rle++;
        *regen_out = regen;
        return (long) (hdr + 1);
    }
    /* Compressed (2) or Treeless (3) */
    if (sf == 0 || sf == 1)
    {
        uint32_t v;

        if (n < 3)
            return -1;
        v = src[0] | ((uint32_t) src[1] << 8) | ((uint32_t) src[2] << 16);
        hdr = 3;
        regen = (v >> 4) & 0x3FF;
        csize = (v >> 14) & 0x3FF;
        streams = sf == 0 ? 1 : 4;
    }
    else if (sf == 2)
    {
        uint32_t v;

        if (n < 4)
            return -1;
        v = rd32(src);
        hdr = 4;
        regen = (v >> 4) & 0x3FFF;
        csize = (v >> 18) & 0x3FFF;
        streams = 4;
    }
    else
    {
        uint64_t v;

        if (n < 5)
            return -1;
        v = rd32(src) | ((uint64_t) src[4] << 32);
        hdr = 5;
        regen = (v >> 4) & 0x3FFFF;
        csize = (v >> 22) & 0x3FFFF;
        streams = 4;
    }
    if (regen > sizeof(z->lit) || hdr + csize > n)
        return -1;
    {
        const uint8_t *p = src + hdr;
        size_t         left = csize;

        if (type == 2)
        {
            long used = huf_read_table(p, left, &z->huf, st);

            if (used < 0)
                return -1;
            z->huf_valid = 1;
            p += used;
            left -= used;
        }
        else if (!z->huf_valid)
            return -1;
        if (streams == 1)
        {
            if (huf_decode_stream(&z->huf, p, left, z->lit, regen) < 0)
                return -1;
        }
        else
        {
            size_t s1, s2, s3, s4, seg = (regen + 3) / 4;

            if (left < 6)
                return -1;
            s1 = p[0] | (p[1] << 8);
            s2 = p[2] | (p[3] << 8);
            s3 = p[4] | (p[5] << 8);
            if (6 + s1 + s2 + s3 > left || seg * 3 > regen)
                return -1;
            s4 = left - 6 - s1 - s2 - s3;
            p += 6;
            if (huf_decode_stream(&z->huf, p, s1, z->lit, seg) < 0 ||
                huf_decode_stream(&z->huf, p + s1, s2, z->lit + seg, seg) < 0 ||
                huf_decode_stream(&z->huf, p + s1 + s2, s3, z->lit + 2 * seg, seg) < 0 ||
                huf_decode_stream(&z->huf, p + s1 + s2 + s3, s4, z->lit + 3 * seg, regen - 3 * seg) < 0)
                return -1;
        }
        if (st)
        {
            if (type == 2)
                (streams == 1 ? st->lit_huf1++ : st->lit_huf4++);
            else
                (streams == 1 ? st->lit_treeless1++ : st->lit_treeless4++);
        }
    }
    *regen_out = regen;
    return (long) (hdr + csize);
}

/* set up one of the three sequence tables; returns bytes consumed or <0 */
static long
seq_table(int mode, const uint8_t *src, size_t n, fse_table *t, int *valid,
          const int16_t *def, int def_n, int def_log, int max_log, int max_sym,
          cryo_oracle_stats *st)
{
    switch (mode)
    {
        case 0:
            fse_build(t, def, def_n, def_log);
            *valid = 1;
            if (st)
                st->mode_predef++;
            return 0;
        case 1:
            if (n < 1 || src[0] > max_sym)
                return -1;
            fse_build_rle(t, src[0]);
            *valid = 1;
            if (st)
                st->mode_rle++;
            return 1;
        case 2:
        {
            int16_t counts[64];
            int     nsym, log;
            long    used = fse_read_counts(src, n, max_log, max_sym, counts, &nsym, &log);

            if (used < 0)
                return -1;
            fse_build(t, counts, nsym, log);
            *valid = 1;
            if (st)
                st->mode_fse++;
            return used;
        }
        default:
            if (!*valid)
                return -1;
            if (st)
                st->mode_repeat++;
            return 0;
    }
}

static long
decode_block(zstd_ctx *z, const uint8_t *src, size_t n, uint8_t *dst, size_t dst_cap,
             size_t op, size_t frame_start, size_t block_max, cryo_oracle_stats *st,
             cryo_oracle_seq_cb cb, void *cb_ctx)
{
    size_t  nlit = 0, lit_pos = 0, start = op;
    long    used = decode_literals(z, src, n, &nlit, st);
    const uint8_t *p;
    size_t  left;
    uint32_t nseq;

    if (used < 0)
        return CRYO_ORACLE_ERR_FORMAT;
    p = src + used;
    left = n - used;
    if (left < 1)
        return CRYO_ORACLE_ERR_INPUT;
    if (p[0] < 128)
    {
        nseq = p[0];
        p += 1;
        left -= 1;
    }
    else if (p[0] < 255)
    {
        if (left < 2)
            return CRYO_ORACLE_ERR_INPUT;
        nseq = ((uint32_t) (p[0] - 128) << 8) + p[1];
        p += 2;
        left -= 2;
    }
    else
    {
        if (left < 3)
            return CRYO_ORACLE_ERR_INPUT;
        nseq = p[1] + ((uint32_t) p[2] << 8) + 0x7F00;
        p += 3;
        left -= 3;
    }
    if (nseq > 0)
    {
        int     modes;
        long    u;
        bbits   bb;
        uint32_t sl, so, sm;

        if (left < 1)
            return CRYO_ORACLE_ERR_INPUT;
        modes = p[0];
        if (modes & 3)
            return CRYO_ORACLE_ERR_FORMAT;
        p++;
        left--;
        u = seq_table((modes >> 6) & 3, p, left, &z->ll, &z->ll_valid, LL_DEFAULT, 36, 6, 9, 35, st);
        if (u < 0)
            return CRYO_ORACLE_ERR_FORMAT;
        p += u;
        left -= u;
        u = seq_table((modes >> 4) & 3, p, left, &z->of, &z->of_valid, OF_DEFAULT, 29, 5, 8, 31, st);
        if (u < 0)
            return CRYO_ORACLE_ERR_FORMAT;
        p += u;
        left -= u;
        u = seq_table((modes >> 2) & 3, p, left, &z->ml, &z->ml_valid, ML_DEFAULT, 53, 6, 9, 52, st);
        if (u < 0)
            return CRYO_ORACLE_ERR_FORMAT;
        p += u;
        left -= u;
        if (binit(&bb, p, left) < 0)
            return CRYO_ORACLE_ERR_FORMAT;
        sl = bread(&bb, z->ll.log);
        so = bread(&bb, z->of.log);
        sm = bread(&bb, z->ml.log);
        for (uint32_t i = 0; i < nseq; i++)
        {
            int      oc = z->of.cell[so].symbol;
            int      mc = z->ml.cell[sm].symbol;
            int      lc = z->ll.cell[sl].symbol;
            uint32_t ov, ml, ll, off;

            if (lc > 35 || mc > 52 || oc > 31)
                return CRYO_ORACLE_ERR_FORMAT;
            ov = (1u << oc) + bread(&bb, oc);
            ml = ML_BASE[mc] + bread(&bb, ML_BITS[mc]);
            ll = LL_BASE[lc] + bread(&bb, LL_BITS[lc]);
            if (ov > 3)
            {
                off = ov - 3;
                z->rep[2] = z->rep[1];
                z->rep[1] = z->rep[0];
                z->rep[0] = off;
            }
            else
            {
                uint32_t idx = ov - 1 + (ll == 0);

                if (st)
                    st->rep_offsets++;
                if (idx == 0)
                    off = z->rep[0];
                else
                {
                    off = idx == 3 ? z->rep[0] - 1 : z->rep[idx];
                    if (idx > 1)
                        z->rep[2] = z->rep[1];
                    z->rep[1] = z->rep[0];
                    z->rep[0] = off;
                }
            }
            if (i + 1 < nseq)
            {
                sl = z->ll.cell[sl].base + bread(&bb, z->ll.cell[sl].nbits);
                sm = z->ml.cell[sm].base + bread(&bb, z->ml.cell[sm].nbits);
                so = z->of.cell[so].base + bread(&bb, z->of.cell[so].nbits);
            }
            if (bb.bit < 0)
                return CRYO_ORACLE_ERR_INPUT;
            /* execute */
            if (lit_pos + ll > nlit)
                return CRYO_ORACLE_ERR_FORMAT;
            if (op + ll + ml > dst_cap)
                return CRYO_ORACLE_ERR_OUTPUT;
            if (op + ll + ml - start > block_max)
                return CRYO_ORACLE_ERR_FORMAT;
            memcpy(dst + op, z->lit + lit_pos, ll);
            lit_pos += ll;
            op += ll;
            if (off == 0 || off > op - frame_start)
                return CRYO_ORACLE_ERR_OFFSET;
            for (uint32_t k = 0; k < ml; k++)
                dst[op + k] = dst[op + k - off];
            op += ml;
            if (st)
            {
                st->sequences++;
                st->literal_bytes += ll;
                st->match_bytes += ml;
                if (ml > st->longest_match)
                    st->longest_match = ml;
                if (off < ml)
                    st->overlap_match_bytes += ml;
                if (ll > st->max_literal_run)
                    st->max_literal_run = ll;
                if (off > st->max_offset)
                    st->max_offset = off;
            }
            if (cb)
                cb(cb_ctx, ll, ml, off);
        }
        if (bb.bit != 0)
            return CRYO_ORACLE_ERR_INPUT;   /* sequence bitstream not consumed exactly */
    }
    else if (left != 0)
        return CRYO_ORACLE_ERR_INPUT;
    /* trailing literals */
    {
        size_t rest = nlit - lit_pos;

        if (op + rest > dst_cap)
            return CRYO_ORACLE_ERR_OUTPUT;
        if (op + rest - start > block_max)
            return CRYO_ORACLE_ERR_FORMAT;
        memcpy(dst + op, z->lit + lit_pos, rest);
        op += rest;
        if (st)
        {
            st->literal_bytes += (uint32_t) rest;
            if (rest > st->max_literal_run)
                st->max_literal_run = (uint32_t) rest;
        }
        if (cb && rest)
            cb(cb_ctx, (uint32_t) rest, 0, 0);
    }
    return (long) op;
}

static zstd_ctx zctx_pool;      /* callers are single-threaded tests; see cryo_oracle_zstd_decode */

long
cryo_oracle_zstd_decode(const uint8_t *src, size_t src_size, uint8_t *dst, size_t dst_cap,
                        cryo_oracle_stats *st, cryo_oracle_seq_cb cb, void *cb_ctx)
{
    size_t      ip = 0, op = 0;
    zstd_ctx   *z = &zctx_pool;

    if (st)
        memset(st, 0, sizeof(*st));
    /* src_size == 0: ZSTD_decompress returns 0 (no frames), not an error */
    while (ip < src_size)
    {
        uint32_t magic;
        int      fhd, fcs_flag, single, checksum, dict_flag;
        uint64_t fcs = 0, window = 0;
        int      have_fcs = 0;
        size_t   frame_start = op, block_max;
        static const int fcs_bytes[4] = {0, 2, 4, 8};
        static const int dict_bytes[4] = {0, 1, 2, 4};

        if (ip + 4 > src_size)
            ERR(CRYO_ORACLE_ERR_INPUT);
        magic = rd32(src + ip);
        if ((magic & 0xFFFFFFF0u) == 0x184D2A50u)
        {
            /* skippable frame */
            if (ip + 8 > src_size)
                ERR(CRYO_ORACLE_ERR_INPUT);
            size_t len = rd32(src + ip + 4);

            if (ip + 8 + len > src_size)
                ERR(CRYO_ORACLE_ERR_INPUT);
            ip += 8 + len;
            continue;
        }
        if (magic != 0xFD2FB528u)
            ERR(CRYO_ORACLE_ERR_FORMAT);
        ip += 4;
        if (ip + 1 > src_size)
            ERR(CRYO_ORACLE_ERR_INPUT);
        fhd = src[ip++];
        fcs_flag = fhd >> 6;
        single = (fhd >> 5) & 1;
        checksum = (fhd >> 2) & 1;
        dict_flag = fhd & 3;
        if (fhd & 0x08)
            ERR(CRYO_ORACLE_ERR_FORMAT);                /* reserved bit */
        if (!single)
        {
            int b, wl;

            if (ip + 1 > src_size)
                ERR(CRYO_ORACLE_ERR_INPUT);
            b = src[ip++];
            wl = 10 + (b >> 3);
            window = (1ULL << wl) + ((1ULL << wl) / 8) * (b & 7);
            if (wl > 27)
                ERR(CRYO_ORACLE_ERR_FORMAT);            /* ZSTD_decompress' default windowLogMax */
        }
        if (ip + dict_bytes[dict_flag] > src_size)
            ERR(CRYO_ORACLE_ERR_INPUT);
        if (dict_flag)
        {
            uint32_t id = 0;

            for (int i = 0; i < dict_bytes[dict_flag]; i++)
                id |= (uint32_t) src[ip + i] << (8 * i);
            ip += dict_bytes[dict_flag];
            if (id != 0)
                ERR(CRYO_ORACLE_ERR_FORMAT);            /* no dictionary on this path */
        }
        {
            int nb = fcs_bytes[fcs_flag];

            if (fcs_flag == 0 && single)
                nb = 1;
            if (ip + nb > src_size)
                ERR(CRYO_ORACLE_ERR_INPUT);
            for (int i = 0; i < nb; i++)
                fcs |= (uint64_t) src[ip + i] << (8 * i);
            if (nb == 2)
                fcs += 256;
            have_fcs = nb > 0;
            ip += nb;
        }
        if (single)
            window = fcs;
        /* RFC 8878: Block_Maximum_Size = min(Window_Size, 128 KiB).  libzstd 1.5.5's one-shot
         * ZSTD_decompress only enforces the 128 KiB constant, so tiny single-segment frames
         * whose blocks exceed their own content size decode fine there; follow the library. */
        block_max = 1 << 17;
        (void) window;
        if (st)
        {
            st->frames++;
            st->window_size = (uint32_t) window;
            st->single_segment = single;
        }
        z->huf_valid = z->ll_valid = z->of_valid = z->ml_valid = 0;
        z->rep[0] = 1;
        z->rep[1] = 4;
        z->rep[2] = 8;
        for (;;)
        {
            uint32_t bh, last, type, size;

            if (ip + 3 > src_size)
                ERR(CRYO_ORACLE_ERR_INPUT);
            bh = src[ip] | ((uint32_t) src[ip + 1] << 8) | ((uint32_t) src[ip + 2] << 16);
            ip += 3;
            last = bh & 1;
            type = (bh >> 1) & 3;
            size = bh >> 3;
            if (type == 3)
                ERR(CRYO_ORACLE_ERR_FORMAT);
            if (type == 0)
            {
                if (size > block_max)
                    ERR(CRYO_ORACLE_ERR_FORMAT);
                if (ip + size > src_size)
                    ERR(CRYO_ORACLE_ERR_INPUT);
                if (op + size > dst_cap)
                    ERR(CRYO_ORACLE_ERR_OUTPUT);
                memcpy(dst + op, src + ip, size);
                ip += size;
                op += size;
                if (st)
                    st->blocks_raw++;
            }
            else if (type == 1)
            {
                if (size > block_max)
                    ERR(CRYO_ORACLE_ERR_FORMAT);
                if (ip + 1 > src_size)
                    ERR(CRYO_ORACLE_ERR_INPUT);
                if (op + size > dst_cap)
                    ERR(CRYO_ORACLE_ERR_OUTPUT);
                memset(dst + op, src[ip], size);
                ip += 1;
                op += size;
                if (st)
                    st->blocks_rle++;
            }
            else
            {
                long r;

                if (size > block_max || size == 0)
                    ERR(CRYO_ORACLE_ERR_FORMAT);
                if (ip + size > src_size)
                    ERR(CRYO_ORACLE_ERR_INPUT);
                r = decode_block(z, src + ip, size, dst, dst_cap, op, frame_start, block_max,
                                 st, cb, cb_ctx);
                if (r < 0)
                    return r;
                op = (size_t) r;
                ip += size;
                if (st)
                    st->blocks_compressed++;
            }
            if (last)
                break;
        }
        if (have_fcs && op - frame_start != fcs)
            ERR(CRYO_ORACLE_ERR_SIZE);
        if (checksum)
        {
            if (ip + 4 > src_size)
                ERR(CRYO_ORACLE_ERR_INPUT);
            if (rd32(src + ip) != (uint32_t) cryo_oracle_xxh64(dst + frame_start, op - frame_start, 0))
                ERR(CRYO_ORACLE_ERR_FORMAT);
            ip += 4;
        }
    }
    return (long) op;
}
