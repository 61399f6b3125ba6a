/*
 * lz4_decode_c.cuh -- LZ4 block decompression, ONE CTA per cryo block (the latency path, and
 * the path for match-rich blocks; lz4_decode_w.cuh is the one-warp-per-block throughput path).
 *
 * Replaces LZ4_decompress_safe as called at reference compression.c:84, same acceptance rules
 * (SURVEY.md D.1).  An LZ4 block is one chain of tokens, and where token i+1 starts is only
 * known after token i has been read: the format has no index.  Walking that chain with one
 * thread costs a dependent shared-memory read per sequence (120 K of them in a dense
 * low-cardinality block); here the chain is found by all threads at once:
 *
 *   stage   the next LZ4C_REGION bytes of the stream -> shared memory, one bulk copy by the
 *           copy engine (cp.async.bulk global -> shared, completion on an mbarrier)
 *   walk    lane l starts at byte 60 * (l - 1) of the region AS IF a token started there and walks to
 *           the end of its own 60-byte segment, recording the token starts it visits there (a bitmap).
 *           A chain that starts at a wrong byte meets the true chain after a few tokens and
 *           then follows it (the token stream self-synchronises), so most of every lane's walk
 *           is the true chain;
 *   link    every lane keeps walking past its segment until it lands on a token start the
 *           owning lane has recorded: from there on the two chains are one;
 *   mark    lane 0's start is a true token start; the lanes reachable from it through the
 *           links are the ones whose chains are true (pointer jumping, log2(lanes) rounds), and
 *           each learns where the true chain enters its segment;
 *   emit    prefix sum of the token counts, then every marked lane decodes its tokens again
 *           and writes (literal length, match length, offset) records, in stream order.
 *
 * The records then go through the CTA-cooperative executor (cryo_cx.cuh) 1 024 at a time;
 * where each sequence's literals are follows from a prefix sum over the records, because the
 * length encoding of LZ4 is unique.  Literals are read from the staged region.
 */
#pragma once
#include "cryo_cx.cuh"

/*
 * Bytes of stream per lane.  Not 64: the lanes of a warp start their walks one segment apart, and with
 * 64-byte segments every one of their shared-memory reads fell on the same two banks (16-way conflicts
 * for as long as the lanes advance in step -- inside the 4 030 length-extension bytes of a sparse block's
 * zero run, for good: that walk took 245 K cycles).  15 words is odd: the starts of 32 consecutive
 * segments fall on 32 different banks.  (The visited-token bitmap of a segment is one 64-bit word.)
 */
#define LZ4C_SEG        60u
#define LZ4C_LANES      (CX_THREADS < 512u ? CX_THREADS : 512u)
#define LZ4C_REGION     (LZ4C_SEG * LZ4C_LANES)         /* stream bytes whose tokens one round parses */
#define LZ4C_STAGE      (LZ4C_REGION + 1024u)           /* staged bytes: the region, its 16-byte alignment, look-ahead */
#define LZ4C_MAXHOPS    24u
#define LZ4C_SERIAL     64u             /* tokens one thread walks before the region is parsed speculatively */
#define LZ4C_SEQCAP     (LZ4C_REGION / 3u + 64u)        /* a sequence that is not the last takes >= 3 bytes */
#define LZ4C_MAXCAP     ((1u << 22) - 1u)               /* record fields are 22 bits */
#define LZ4C_BAD        0xFFFFFFFFu
#define LZ4C_END        LZ4C_LANES                      /* sentinel lane of the links */

/* link kinds */
#define LZ4C_MERGED 0u
#define LZ4C_EXIT   1u
#define LZ4C_DEAD   2u              /* the chain runs into a malformed token */
#define LZ4C_IRREG  3u              /* no merge within LZ4C_MAXHOPS tokens */

/* shared memory of one CTA */
#define LZ4C_OFF_RING   0u
#define LZ4C_OFF_STAGE  (LZ4C_OFF_RING + CX_RING)
#define LZ4C_OFF_PAT    (LZ4C_OFF_STAGE + LZ4C_STAGE)
#define LZ4C_OFF_DL     (LZ4C_OFF_PAT + CX_PAT + 32u)
#define LZ4C_OFF_CXSH   (LZ4C_OFF_DL + 2u * CX_SPAN)
#define LZ4C_OFF_PARSE  ((LZ4C_OFF_CXSH + (uint32_t) sizeof(CxSh) + 15u) & ~15u)
struct Lz4cParse
{
    unsigned long long mbar;                    /* completion of the bulk copy */
    unsigned long long vis[LZ4C_LANES];         /* token starts the lane's chain visits in its segment */
    uint32_t    exitp[LZ4C_LANES];              /* first position of the chain at or past the segment end */
    uint32_t    link[LZ4C_LANES];               /* where the chain merges / leaves the region */
    uint32_t    extc[LZ4C_LANES];               /* tokens between exitp and link | kind << 24 */
    uint32_t    entry[LZ4C_LANES + 1];          /* where the true chain enters the segment */
    uint16_t    jmp[2][LZ4C_LANES + 2];
    uint8_t     mark[LZ4C_LANES + 4];
    uint32_t    fix;                            /* a marked lane has an irregular link */
    uint32_t    stop;                           /* kind of the last link of the true chain, its position */
    uint32_t    stop_pos;
    uint32_t    bad;                            /* emit found a malformed token on the true chain */
    uint32_t    saw_last;
    uint32_t    ser_cnt, ser_pos, ser_flags;    /* the serial walk: records, where it stopped, 1: malformed 2: last sequence seen 4: it covered the region */
    uint32_t    ffmap[(LZ4C_STAGE / 16u + 31u) / 32u];  /* bit g: the 16 staged bytes of group g are all 0xFF (and before the stream's end) */
};
#define LZ4C_SMEM       (LZ4C_OFF_PARSE + (uint32_t) sizeof(Lz4cParse))
static_assert(LZ4C_SMEM <= 232448u, "one CTA's shared memory on sm_100");

struct Lz4cIn
{
    const uint8_t *base;        /* 16-byte aligned address at or before the stream */
    const uint8_t *stage;       /* shared: base[sbase, sbase + slen) */
    const uint32_t *ffmap;      /* shared: which 16-byte groups of the stage are all 0xFF (Lz4cParse) */
    uint32_t    sbase, slen;
    uint32_t    end;            /* stream end in `base` coordinates */
};

CRYO_DEV uint32_t lz4c_byte(const Lz4cIn &in, uint32_t p)
{
    const uint32_t r = p - in.sbase;

    return r < in.slen ? in.stage[r] : in.base[p];
}

struct Lz4cTok
{
    uint32_t ll, ml, off, lit, next;
    uint32_t st;                /* 0: sequence with a match, 1: the last sequence (literals only), 2: malformed */
};

/* the 8 stream bytes at p, little-endian; bytes at or past in.end read as zero */
CRYO_DEV unsigned long long lz4c_ld8(const Lz4cIn &in, uint32_t p)
{
    const uint32_t r = p - in.sbase;

    if (r + 12u <= in.slen && p + 8u <= in.end)
    {
        const uint32_t *w = reinterpret_cast<const uint32_t *>(in.stage) + (r >> 2);
        const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], sh = (r & 3u) * 8u;

        return (unsigned long long) __funnelshift_r(w0, w1, sh) | ((unsigned long long) __funnelshift_r(w1, w2, sh) << 32);
    }
    unsigned long long v = 0;

    for (uint32_t i = 0; i < 8u && p + i < in.end; i++)
        v |= (unsigned long long) lz4c_byte(in, p + i) << (8u * i);
    return v;
}

/*
 * Length extension at q: bytes are added until one is below 255.  Sixteen bytes per step (the
 * zero run of a sparse cryo block is a match of ~1 MB: 4 030 extension bytes).  False: the
 * stream ends inside the run.
 */
CRYO_DEV bool lz4c_extension(const Lz4cIn &in, uint32_t &q, uint32_t &len)
{
    for (;;)
    {
        if (q >= in.end)
            return false;
        /* whole groups of 0xFF at once from the map: a lane alone takes ~300 cycles per 16-byte step below (a chain of
         * dependent instructions), and every lane that starts inside a long run scans to its end */
        {
            const uint32_t r = q - in.sbase;

            if ((r & 15u) == 0u && r < in.slen)
            {
                const uint32_t g = r >> 4, w = ~(in.ffmap[g >> 5] >> (g & 31u));
                uint32_t       n = w ? (uint32_t) __ffs((int) w) - 1u : 32u;

                if (n > 32u - (g & 31u))
                    n = 32u - (g & 31u);
                if (n)
                {
                    len += n * 16u * 255u;
                    q += 16u * n;
                    if (len > (1u << 24))
                        return false;
                    continue;
                }
            }
        }
        const unsigned long long a = lz4c_ld8(in, q), b = lz4c_ld8(in, q + 8u);

        if (a == ~0ull && b == ~0ull && q + 16u <= in.end)
        {
            /* sixteen more; or only up to the next group boundary of the stage, so that the map takes over */
            const uint32_t mis = (q - in.sbase) & 15u, step = mis && q - in.sbase < in.slen ? 16u - mis : 16u;

            len += step * 255u;
            q += step;
            if (len > (1u << 24))
                return false;
            continue;
        }
        const unsigned long long x = a == ~0ull ? b : a;
        const uint32_t base = a == ~0ull ? 8u : 0u;
        const uint32_t k = x == ~0ull ? 8u : (uint32_t) (__ffsll((long long) ~x) - 1) >> 3;     /* first byte below 255 */

        if (k == 8u)
        {
            /* (only when the stream ends within these 16 bytes) */
            len += (base + 8u) * 255u;
            q += base + 8u;
            continue;
        }
        if (q + base + k >= in.end)
            return false;
        len += (base + k) * 255u + (uint32_t) ((x >> (8u * k)) & 0xFFu);
        q += base + k + 1u;
        return true;
    }
}

/* decode the token at p (p < in.end) */
CRYO_DEV void lz4c_tok(const Lz4cIn &in, uint32_t p, Lz4cTok &t)
{
    const unsigned long long W = lz4c_ld8(in, p);
    const uint32_t tk = (uint32_t) W & 0xFFu;
    uint32_t q = p + 1, ll = tk >> 4, ml = tk & 15u;

    t.st = 2;
    t.ml = 0;
    t.off = 0;
    t.next = LZ4C_BAD;
    if (ll == 15u && !lz4c_extension(in, q, ll))
        return;
    t.ll = ll;
    t.lit = q;
    if (ll > in.end - q)
        return;
    q += ll;
    if (q == in.end)
    {
        t.st = 1;
        t.next = in.end;
        return;
    }
    if (q + 2u > in.end)
        return;
    /* the offset: still inside the 8 bytes read for the token when the literal run is short */
    if (q - p <= 6u)
        t.off = (uint32_t) (W >> (8u * (q - p))) & 0xFFFFu;
    else
        t.off = (uint32_t) lz4c_ld8(in, q) & 0xFFFFu;
    q += 2;
    if (ml == 15u && !lz4c_extension(in, q, ml))
        return;
    t.ml = ml + 4u;
    t.next = q;
    t.st = 0;
}

/*
 * The same for the one thread of the serial walk: byte loads, no 64-bit assembly -- a lone thread runs one
 * dependent instruction every 5-6 cycles, and lz4c_tok's 150 instructions per literal-heavy token came to 800.
 */
CRYO_DEV void lz4c_tok_serial(const Lz4cIn &in, uint32_t p, Lz4cTok &t)
{
    const uint32_t tk = lz4c_byte(in, p);
    uint32_t q = p + 1, ll = tk >> 4, ml = tk & 15u;

    t.st = 2;
    t.ml = 0;
    t.off = 0;
    t.next = LZ4C_BAD;
    if (ll == 15u)
        for (;;)
        {
            /* at a group boundary of the stage the map-driven scan takes over (long runs of 0xFF) */
            if (((q - in.sbase) & 15u) == 0u)
            {
                if (!lz4c_extension(in, q, ll))
                    return;
                break;
            }
            if (q >= in.end)
                return;
            const uint32_t x = lz4c_byte(in, q++);

            ll += x;
            if (x != 255u)
                break;
            if (ll > (1u << 24))
                return;
        }
    t.ll = ll;
    t.lit = q;
    if (ll > in.end - q)
        return;
    q += ll;
    if (q == in.end)
    {
        t.st = 1;
        t.next = in.end;
        return;
    }
    if (q + 2u > in.end)
        return;
    t.off = lz4c_byte(in, q) | (lz4c_byte(in, q + 1u) << 8);
    q += 2;
    if (ml == 15u)
        for (;;)
        {
            if (((q - in.sbase) & 15u) == 0u)
            {
                if (!lz4c_extension(in, q, ml))
                    return;
                break;
            }
            if (q >= in.end)
                return;
            const uint32_t x = lz4c_byte(in, q++);

            ml += x;
            if (x != 255u)
                break;
            if (ml > (1u << 24))
                return;
        }
    t.ml = ml + 4u;
    t.next = q;
    t.st = 0;
}

/* bytes of length extension that encode a length field of value v (the encoding is unique) */
CRYO_DEV uint32_t lz4c_ext(uint32_t v)
{
    return v < 15u ? 0u : (v - 15u) / 255u + 1u;
}

/* lane l walks to the end of its segment from position p0 (inside it, or before it: the run-up):
 * token starts visited inside the segment, and the exit position */
CRYO_DEV void lz4c_walk(const Lz4cIn &in, uint32_t segstart, uint32_t segend, uint32_t p0,
                        unsigned long long &vis, uint32_t &exitp)
{
    uint32_t p = p0;
    unsigned long long v = 0;

    while (p < segend)
    {
        Lz4cTok t;

        if (p >= segstart)
            v |= 1ull << (p - segstart);
        lz4c_tok(in, p, t);
        p = t.next;             /* LZ4C_BAD ends the loop */
    }
    vis = v;
    exitp = p;
}

/* from the lane's exit position on, until the chain merges with the chain of a later lane */
CRYO_DEV void lz4c_extend(const Lz4cIn &in, const Lz4cParse *ps, uint32_t rp, uint32_t re, uint32_t q,
                          uint32_t &link, uint32_t &extc)
{
    uint32_t cnt = 0, kind;

    for (;;)
    {
        if (q == LZ4C_BAD)
        {
            kind = LZ4C_DEAD;
            break;
        }
        if (q >= re)
        {
            kind = LZ4C_EXIT;
            break;
        }
        const uint32_t o = (q - rp) / LZ4C_SEG;

        if ((ps->vis[o] >> ((q - rp) % LZ4C_SEG)) & 1ull)
        {
            kind = LZ4C_MERGED;
            break;
        }
        if (cnt >= LZ4C_MAXHOPS)
        {
            kind = LZ4C_IRREG;
            break;
        }
        Lz4cTok t;

        lz4c_tok(in, q, t);
        cnt++;
        q = t.next;
    }
    link = q;
    extc = cnt | (kind << 24);
}

/*
 * One round: the tokens that start in [rp, re) -> records at gseq[0, nseq).  rp is a true token
 * start.  Returns the number of records; next_rp is where the chain continues (>= re), st the
 * block status if the chain is malformed, saw_last whether the last sequence of the block was
 * among them.
 */
CRYO_DEV uint32_t lz4c_parse_round(const Lz4cIn &in, Lz4cParse *ps, CxSh *sh, uint32_t rp, uint32_t re,
                                   unsigned long long *gseq, uint32_t &next_rp, int &st, bool &saw_last, uint32_t tid,
                                   bool try_serial)
{
    const uint32_t l = tid;

    /*
     * A region of few tokens -- long literal runs, the zero run of a sparse block -- is crossed by the true chain in a
     * few hops, while the lanes of the speculative walk hop through literal bytes and merge nowhere.  When the region
     * before this one was sparse, one thread first walks the chain; if LZ4C_SERIAL tokens do not cover the region the
     * speculative parse takes over (its records overwrite these).  A lone thread needs ~760 cycles per token (a chain
     * of dependent instructions at one per ~6 cycles), which is what the speculative parse costs per token at 55
     * tokens per region: the serial walk only wins below that, and removes the repairs of irregular links there.
     */
    if (try_serial)
    {
        if (tid == 0)
        {
            uint32_t p = rp, cnt = 0, flags = 0;

            while (p < re && cnt < LZ4C_SERIAL)
            {
                Lz4cTok t;

                lz4c_tok_serial(in, p, t);
                if (t.st == 2 || t.ll > LZ4C_MAXCAP || t.ml > LZ4C_MAXCAP)
                {
                    flags |= 1u;
                    break;
                }
                gseq[cnt++] = (unsigned long long) t.ll | ((unsigned long long) t.ml << 22) |
                              ((unsigned long long) t.off << 44) | ((unsigned long long) (t.st == 1) << 63);
                if (t.st == 1)
                    flags |= 2u;
                p = t.next;
                if (cnt == 16u && p - rp < 512u)
                    break;              /* dense after all: 16 tokens in less than 512 bytes */
            }
            if (p >= re || (flags & 1u))
                flags |= 4u;
            ps->ser_cnt = cnt;
            ps->ser_pos = p;
            ps->ser_flags = flags;
        }
        __syncthreads();
        const uint32_t flags = ps->ser_flags;

        if (flags & 4u)
        {
            st = (flags & 1u) ? ST_INPUT : ST_OK;
            saw_last = (flags & 2u) != 0;
            next_rp = ps->ser_pos;
            return ps->ser_cnt;
        }
    }
    const bool     lane_on = l < LZ4C_LANES && rp + l * LZ4C_SEG < re;
    const uint32_t segstart = rp + l * LZ4C_SEG;
    const uint32_t segend = segstart + LZ4C_SEG < re ? segstart + LZ4C_SEG : re;

    /* walk */
    if (l < LZ4C_LANES)
    {
        unsigned long long v = 0;
        uint32_t x = LZ4C_BAD;

        /* the walk starts one segment early (not before rp, which is a true token start): by the time
         * the chain enters its own segment it has almost always fallen in with the true chain, so the
         * true chain of the lane before finds it at once (without the run-up one lane in forty of a dense
         * block missed it within LZ4C_MAXHOPS tokens and had to be repaired, one serial walk each) */
        if (lane_on)
            lz4c_walk(in, segstart, segend, l == 0 ? segstart : segstart - LZ4C_SEG, v, x);
        ps->vis[l] = v;
        ps->exitp[l] = x;
    }
    if (tid == 0)
    {
        ps->bad = 0;
        ps->saw_last = 0;
    }
    __syncthreads();
    CXP(1)
    /* link */
    if (lane_on)
        lz4c_extend(in, ps, rp, re, ps->exitp[l], ps->link[l], ps->extc[l]);
    __syncthreads();
    CXP(2)
    /* mark (and repair an irregular link of a marked lane, then mark again) */
    for (;;)
    {
        CXP_COUNT(19, 1)
        if (l < LZ4C_LANES)
        {
            uint32_t nx = LZ4C_END;

            if (lane_on && (ps->extc[l] >> 24) == LZ4C_MERGED)
                nx = (ps->link[l] - rp) / LZ4C_SEG;
            ps->jmp[0][l] = (uint16_t) nx;
            ps->mark[l] = l == 0 ? 1 : 0;
            ps->entry[l] = LZ4C_BAD;
        }
        if (tid == 0)
        {
            ps->jmp[0][LZ4C_END] = (uint16_t) LZ4C_END;
            ps->jmp[1][LZ4C_END] = (uint16_t) LZ4C_END;
            ps->mark[LZ4C_END] = 0;
            ps->fix = 0;
            ps->entry[0] = rp;
        }
        __syncthreads();
        {
            uint32_t cur = 0;

            for (uint32_t span = 1; span < LZ4C_LANES; span <<= 1, cur ^= 1u)
            {
                if (l < LZ4C_LANES)
                {
                    const uint32_t j = ps->jmp[cur][l];

                    if (ps->mark[l])
                        ps->mark[j] = 1;
                    ps->jmp[cur ^ 1u][l] = ps->jmp[cur][j];
                }
                __syncthreads();
            }
        }
        /* entries, and the end of the true chain */
        if (lane_on && ps->mark[l])
        {
            const uint32_t kind = ps->extc[l] >> 24;

            if (kind == LZ4C_MERGED)
                ps->entry[(ps->link[l] - rp) / LZ4C_SEG] = ps->link[l];
            else
            {
                ps->stop = kind;
                ps->stop_pos = ps->link[l];
                if (kind == LZ4C_IRREG)
                    ps->fix = 1u + l;
            }
        }
        __syncthreads();
        if (ps->fix == 0)
            break;
        /* the true chain reaches link[f] without meeting the owner's chain: the owner walks again
         * from there (its old chain was not the true one), and links again */
        {
            const uint32_t f = ps->fix - 1u, q = ps->link[f], o = (q - rp) / LZ4C_SEG;

            __syncthreads();
            if (l == o)
            {
                unsigned long long v;
                uint32_t x;

                lz4c_walk(in, segstart, segend, q, v, x);
                ps->vis[l] = v;
                ps->exitp[l] = x;
            }
            __syncthreads();
            if (l == o)
                lz4c_extend(in, ps, rp, re, ps->exitp[l], ps->link[l], ps->extc[l]);
            if (l == f)
                ps->extc[l] = (ps->extc[l] & 0xFFFFFFu) | (LZ4C_MERGED << 24);
            __syncthreads();
        }
    }
    CXP(3)
    /* emit: count, prefix sum, write */
    const bool     marked = lane_on && ps->mark[l];
    const uint32_t ent = marked ? ps->entry[l] : 0u;
    uint32_t cnt = 0, dummy = 0;

    if (marked)
        cnt = (uint32_t) __popcll(ps->vis[l] >> (ent - segstart)) + (ps->extc[l] & 0xFFFFFFu);
    const uint32_t mycnt = cnt;

    cx_scan2(sh, cnt, dummy, tid);
    const uint32_t total = sh->wsa[CX_WARPS];

    if (marked && total <= LZ4C_SEQCAP)
    {
        uint32_t p = ent, at = cnt - mycnt, wrote = 0;
        const uint32_t stop_at = ps->link[l];

        while (p != stop_at && p < re)
        {
            Lz4cTok t;

            lz4c_tok(in, p, t);
            if (t.st == 2 || t.ll > LZ4C_MAXCAP || t.ml > LZ4C_MAXCAP)
            {
                ps->bad = 1;
                break;
            }
            gseq[at++] = (unsigned long long) t.ll | ((unsigned long long) t.ml << 22) |
                         ((unsigned long long) t.off << 44) | ((unsigned long long) (t.st == 1) << 63);
            wrote++;
            if (t.st == 1)
                ps->saw_last = 1;
            p = t.next;
        }
        if (wrote != mycnt)
            ps->bad = 1;                /* the chain died before the link (malformed token) */
    }
    __syncthreads();
    CXP(4)
    st = ST_OK;
    if (ps->bad || ps->stop == LZ4C_DEAD || total > LZ4C_SEQCAP)
        st = ST_INPUT;
    saw_last = ps->saw_last != 0;
    next_rp = ps->stop_pos;
    return total;
}

/* global scratch per CTA */
#define LZ4C_SCRATCH_BYTES ((size_t) LZ4C_SEQCAP * 8u)

/* one CTA decodes one block; smem: LZ4C_SMEM bytes, gseq: LZ4C_SEQCAP records private to the CTA */
CRYO_DEV void lz4c_decode_block(const uint8_t *src, uint32_t csize, uint8_t *out, uint32_t cap, uint32_t *out_size,
                                int32_t *status, uint8_t *smem, unsigned long long *gseq, uint32_t tid)
{
    Cx         cx;
    Lz4cIn     in;
    CxSh      *sh = reinterpret_cast<CxSh *>(smem + LZ4C_OFF_CXSH);
    Lz4cParse *ps = reinterpret_cast<Lz4cParse *>(smem + LZ4C_OFF_PARSE);
    uint8_t   *stage = smem + LZ4C_OFF_STAGE;
    const uint32_t delta = (uint32_t) ((uintptr_t) src & 15u);
    int        err = ST_OK;
    bool       got_last = false;
    uint32_t   phase = 0;

    cx_init(cx, out, cap, smem + LZ4C_OFF_RING, smem + LZ4C_OFF_PAT, reinterpret_cast<uint16_t *>(smem + LZ4C_OFF_DL), sh);
    in.base = src - delta;
    in.end = csize + delta;
    in.stage = stage;
    in.ffmap = ps->ffmap;
    in.sbase = 0;
    in.slen = 0;
    if (csize == 0)
        err = ST_INPUT;
#ifndef CRYO_EMU
    if (tid == 0)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t) __cvta_generic_to_shared(&ps->mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
#endif
    __syncthreads();
#if defined(CX_PROF) && !defined(CRYO_EMU)
    if (threadIdx.x == 0 && blockIdx.x == 0)
        cx_prof_last = clock64();
#endif
    uint32_t rp = delta;
    bool     sparse = true;             /* (try the serial walk on the first region) */

    while (err == ST_OK && rp < in.end)
    {
        /* ---- stage [rp & ~15, ...) through the copy engine ---- */
        const uint32_t lim = (in.end + 15u) & ~15u;

        in.sbase = rp & ~15u;
        in.slen = lim - in.sbase < LZ4C_STAGE ? lim - in.sbase : LZ4C_STAGE;
        __syncthreads();                /* everybody is done with the previous stage contents */
#ifdef CRYO_EMU
        for (uint32_t a = 16u * tid; a < in.slen; a += 16u * CX_THREADS)
            st16(stage + a, ld16(in.base + in.sbase + a));
        __syncthreads();
#else
        if (tid == 0)
        {
            const uint32_t mb = (uint32_t) __cvta_generic_to_shared(&ps->mbar);

            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(in.slen) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"((uint32_t) __cvta_generic_to_shared(stage)), "l"(in.base + in.sbase), "r"(in.slen), "r"(mb)
                         : "memory");
        }
        {
            const uint32_t mb = (uint32_t) __cvta_generic_to_shared(&ps->mbar);
            uint32_t ok = 0;

            while (!ok)
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                             : "=r"(ok) : "r"(mb), "r"(phase) : "memory");
        }
        phase ^= 1u;
#endif
        /* the all-0xFF map of the staged bytes (lz4c_extension) */
        for (uint32_t g0 = 0; g0 < (LZ4C_STAGE / 16u + 31u) / 32u * 32u; g0 += CX_THREADS)
        {
            const uint32_t g = g0 + tid;
            bool           ff = false;

            if (16u * g + 16u <= in.slen && in.sbase + 16u * g + 16u <= in.end)
            {
                const uint4 v = ld16(stage + 16u * g);

                ff = (v.x & v.y & v.z & v.w) == 0xFFFFFFFFu;
            }
            const uint32_t m = __ballot_sync(CRYO_FULL, ff);

            if ((tid & 31u) == 0 && (g >> 5) < (LZ4C_STAGE / 16u + 31u) / 32u)
                ps->ffmap[g >> 5] = m;
        }
        __syncthreads();
        CXP(0)
        const uint32_t re = in.end - rp < LZ4C_REGION ? in.end : rp + LZ4C_REGION;
        uint32_t next_rp = in.end;
        bool     saw_last = false;
        const uint32_t nseq = lz4c_parse_round(in, ps, sh, rp, re, gseq, next_rp, err, saw_last, tid, sparse);

        sparse = nseq <= LZ4C_SERIAL;   /* the next region probably looks like this one */

        if (err != ST_OK)
            break;
        /* ---- execute the records, CX_THREADS at a time ---- */
        uint32_t c0 = 0, tokpos = rp;   /* record c0's token is at tokpos */

        while (c0 < nseq)
        {
            const uint32_t n = nseq - c0 < CX_THREADS ? nseq - c0 : CX_THREADS;
            const bool     valid = tid < n;
            const unsigned long long r = valid ? gseq[c0 + tid] : 0ull;
            const uint32_t ll = (uint32_t) r & 0x3FFFFFu, ml = (uint32_t) (r >> 22) & 0x3FFFFFu;
            const uint32_t off = (uint32_t) (r >> 44) & 0xFFFFu;
            const bool     last = (r >> 63) != 0;
            const uint32_t hdr = 1u + lz4c_ext(ll);
            const uint32_t inlen = valid ? hdr + ll + (last ? 0u : 2u + lz4c_ext(ml - 4u)) : 0u;
            uint32_t cum = valid ? ll + ml : 0u, icum = inlen;

            __syncthreads();            /* scan scratch of the previous chunk / of the parse */
            CXP(5)
            cx_scan2(sh, cum, icum, tid);
            CXP(6)
            const uint32_t lit = tokpos + icum - inlen + hdr;       /* first literal byte, `base` coordinates */
            const uint32_t start = cx.pos + cum - ll - ml;
            int            pre = ST_OK;

            /* LZ4_decompress_safe: a literal run that ends within 12 bytes of the output capacity or
             * within 8 bytes of the input end must be the last one; a match must leave 5 bytes */
            if (valid && !last)
            {
                if (start + ll + 12u > cap || lit + ll + 8u > in.end)
                    pre = (lit + ll + 8u > in.end && start + ll + 12u <= cap) ? ST_INPUT : ST_OUTPUT;
                else if (start + ll + ml + 5u > cap)
                    pre = ST_OUTPUT;
            }
            const uint32_t rl = lit - in.sbase;
            const uint8_t *lp = rl < in.slen && rl + ll <= in.slen ? stage + rl : in.base + lit;
            CXP(7)
            const uint32_t k = cx_chunk(cx, n, ll, ml, off, lp, -1, cum, pre, tid);

            if (k == 0)
            {
                err = cx.err;
                break;
            }
            /* the in-stream length of the k sequences just executed */
            if (tid == k - 1u)
                sh->bcast[0] = icum;
            __syncthreads();
            tokpos += sh->bcast[0];
            c0 += k;
        }
        if (saw_last)
            got_last = true;
        rp = next_rp;
    }
    if (err == ST_OK && !got_last)
        err = ST_INPUT;                 /* the stream ends with a match, or without its last literals */
    cx_finish(cx, tid);
    CXP(14)
    if (tid == 0)
    {
        *out_size = err == ST_OK ? cx.pos : 0u;
        *status = err;
    }
}
