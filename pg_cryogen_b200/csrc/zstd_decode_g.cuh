/*
 * zstd_decode_g.cuh -- batched zstd frame decompression, one GROUP of W lanes per cryo block
 * (cryo_group.cuh): with W = 8 a warp decodes four frames at once.  This is the throughput
 * path behind cryogpu_decompress_*; zstd_decode_w.cuh (one warp per frame) and
 * zstd_decode.cuh (one CTA per frame) are the earlier variants and stay selectable.
 *
 * Replaces ZSTD_decompress as called at reference compression.c:116; same coverage of
 * RFC 8878 and same shared-memory plan per frame as zstd_decode_w.cuh (9 216 bytes: Huffman
 * table / ring + windows, then the three FSE tables).  What changes is who repeats the
 * scalar work.  The FSE state walk, the block and frame headers and the per-sequence
 * bookkeeping are warp-uniform: on one warp per frame all 32 lanes execute them for one
 * frame, and the kernel was bound by instruction issue.  Here eight lanes execute them, so
 * one warp instruction advances four frames; Huffman streams use 4 lanes of every 8 (16 of
 * 32 instead of 4 of 32); copies and fills move W x 16 bytes per instruction per frame.
 */
#pragma once
#include "cryo_gexec.cuh"
#include "zstd_decode_w.cuh"

#define ZSG_W         16
#define ZSG_GROUPS    8                          /* frames per CTA: 2 warps x 4 groups */
#define ZSG_THREADS   (ZSG_W * ZSG_GROUPS)
#define ZSG_CTAS_PER_SM 3
#define ZSG_SMEM      (ZSG_GROUPS * ZSW_PER_WARP)

#if GX_RING != WX_RING
#error "the group executor shares the per-frame shared-memory plan of zstd_decode_w.cuh"
#endif

/*
 * FSE decoding-table build (RFC 8878 4.1.1) by a group of W lanes; same result as the serial
 * fse_build_table: cell = symbol | nbits << 8 | base << 16.  counts[nsym] in shared memory;
 * next[64] and cum[65] are shared scratch.
 */
template <int W>
CRYO_DEV void fse_build_table_g(uint32_t *cell, const int16_t *counts, int nsym, int log,
                                uint16_t *next, uint16_t *cum, const Grp<W> &g)
{
    const uint32_t size = 1u << log, mask = size - 1u;
    const uint32_t step = (size >> 1) + (size >> 3) + 3u;
    const uint32_t lane = g.lane, lt = (1u << lane) - 1u;
    uint32_t nlow = 0;

    /* low-probability symbols sit at the top of the table, one cell each; prefix sums of the
     * positive counts in symbol order (at most 53 symbols: one lane) */
    if (lane == 0)
    {
        uint32_t a = 0;

        for (int s = 0; s < nsym; s++)
        {
            const int c = counts[s];

            cum[s] = (uint16_t) a;
            if (c == -1)
            {
                cell[size - 1u - nlow] = (uint32_t) s;
                nlow++;
                next[s] = 1;
            }
            else
            {
                next[s] = (uint16_t) c;
                a += c > 0 ? (uint32_t) c : 0u;
            }
        }
    }
    nlow = g_shfl(g, nlow, 0);
    const uint32_t high = size - 1u - nlow;         /* last cell of the spread region */
    uint32_t inv = step;                            /* Newton: inv * step == 1 (mod 2^32) */

#pragma unroll
    for (int k = 0; k < 5; k++)
        inv *= 2u - step * inv;
    g_sync(g);
    /* spread: the j-th spread slot belongs to symbol s with cum[s] <= j < cum[s] + count[s];
     * it lands on the j-th visited position that is not a low cell */
    for (int s = 0; s < nsym; s++)
    {
        const int c = counts[s];

        if (c <= 0)
            continue;
        const uint32_t c0 = cum[s];

        for (uint32_t j = c0 + lane; j < c0 + (uint32_t) c; j += W)
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
    g_sync(g);
    /* state numbering in cell order: the r-th cell of symbol s gets next = count[s] + r */
    for (uint32_t p0 = 0; p0 < size; p0 += W)
    {
        const uint32_t p = p0 + lane;
        const uint32_t s = cell[p];
        const uint32_t m = g_match_any(g, s);
        const uint32_t nx = (uint32_t) next[s] + (uint32_t) __popc(m & lt);

        g_sync(g);
        if ((m & lt) == 0)
            next[s] = (uint16_t) (next[s] + __popc(m));
        const uint32_t nb = (uint32_t) (log - zs_highbit(nx));

        cell[p] = s | (nb << 8) | ((((nx << nb) - size) & 0xFFFFu) << 16);
        g_sync(g);
    }
}

/* one sequence table (t: 0 LL, 1 OF, 2 ML); returns bytes of description consumed or ~0u */
template <int W>
CRYO_DEV uint32_t zsg_seq_table(ZswState &z, int mode, int t, const uint8_t *p, uint32_t left,
                                uint32_t *cell, uint8_t *work, const uint32_t *predef, int &logv,
                                const Grp<W> &g)
{
    const uint32_t lane = g.lane;
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

            for (uint32_t i = lane; i < n; i += W)
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
            used = g_shfl(g, (uint32_t) used, 0);
            nsym = g_shfl(g, (uint32_t) nsym, 0);
            log = g_shfl(g, (uint32_t) log, 0);
            if (used == 0)
                return ~0u;
            g_sync(g);
            fse_build_table_g<W>(cell, counts, nsym, log, next, cum, g);
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
    g_sync(g);
    for (uint32_t i = lane; i < (1u << logv); i += W)
    {
        const uint32_t c = cell[i], sym = c & 0xFFu;
        const uint32_t xb = t == 1 ? sym : (t == 0 ? CRYO_GLD(ZS_LL_PACK[sym]) : CRYO_GLD(ZS_ML_PACK[sym])) >> 24;

        cell[i] = (c & 0x03FFFFFFu) | (xb << 26);
    }
    g_sync(g);
    return ret;
}

/*
 * Huffman tree description -> decoding table huf[1 << log] (u16: symbol | nbits << 8), same
 * result as huf_build_table (zstd_decode.cuh) with the per-symbol ranking done by the whole
 * warp: 32 symbols per step, rank inside the step by __match_any_sync, running per-weight
 * counters in shared memory.  Returns bytes used by the description, 0 on error.
 */
template <int W>
CRYO_DEV uint32_t zsg_huf_build(const uint8_t *src, uint32_t n, uint16_t *huf, uint8_t *work,
                                int32_t *log_out, const Grp<W> &g)
{
    const uint32_t lane = g.lane;
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
        for (uint32_t i = lane; i < nw; i += W)
        {
            uint32_t b = src[1 + i / 2];

            weights[i] = (uint8_t) ((i & 1) ? (b & 15u) : (b >> 4));
        }
        g_sync(g);
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
        bad = g_shfl(g, (uint32_t) bad, 0);
        nw = g_shfl(g, (uint32_t) nw, 0);
        if (bad)
            return 0;
        g_sync(g);
    }
    /* sum of 2^(w-1), implied last weight */
    uint32_t sum = 0, over = 0;

    for (uint32_t i = lane; i < nw; i += W)
    {
        uint32_t w = weights[i];

        if (w > 11)
            over = 1;
        else if (w)
            sum += 1u << (w - 1);
    }
    sum = g_reduce_add(g, sum);
    over = g_reduce_or(g, over);
    if (over || sum == 0)
        return 0;
    const int log = zs_highbit(sum) + 1;

    if (log > 11)
        return 0;
    const uint32_t left = (1u << log) - sum;

    if (left & (left - 1))
        return 0;
    if (lane == 0)
        weights[nw] = (uint8_t) (zs_highbit(left) + 1);
    nw += 1;
    for (uint32_t i = lane; i < 32; i += W)
        rankc[i] = 0;
    g_sync(g);
    /* pass 1: symbols per weight */
    const uint32_t lt = (1u << lane) - 1u;

    for (uint32_t s0 = 0; s0 < nw; s0 += W)
    {
        const uint32_t s = s0 + lane;
        const uint32_t w = s < nw ? weights[s] : 0u;
        const uint32_t m = g_match_any(g, w);

        if (w && (m & lt) == 0)
            rankc[w] += (uint32_t) __popc(m);
        g_sync(g);
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
    g_sync(g);
    /* pass 2: first cell of every symbol */
    for (uint32_t s0 = 0; s0 < nw; s0 += W)
    {
        const uint32_t s = s0 + lane;
        const uint32_t w = s < nw ? weights[s] : 0u;
        const uint32_t m = g_match_any(g, w);

        if (w)
            symstart[s] = (uint16_t) (rankc[16 + w] + ((uint32_t) __popc(m & lt) << (w - 1)));
        g_sync(g);
        if (w && (m & lt) == 0)
            rankc[16 + w] += (uint32_t) __popc(m) << (w - 1);
        g_sync(g);
    }
    /* fill: long codes (few cells) one symbol per lane, short codes by the whole warp */
    for (uint32_t s0 = 0; s0 < nw; s0 += W)
    {
        const uint32_t s = s0 + lane;
        const uint32_t w = s < nw ? weights[s] : 0u;
        const uint32_t len = w ? 1u << (w - 1) : 0u;
        const uint32_t st = w ? symstart[s] : 0u;
        const uint16_t ent = (uint16_t) (s | ((uint32_t) (log + 1 - (int) w) << 8));

        if (len && len <= 8)
            for (uint32_t i = 0; i < len; i++)
                huf[st + i] = ent;
        uint32_t big = g_ballot(g, len > 8);

        while (big)
        {
            const int      k = __ffs((int) big) - 1;
            const uint32_t klen = g_shfl(g, len, (uint32_t) k), kst = g_shfl(g, st, (uint32_t) k);
            const uint32_t kent = g_shfl(g, (uint32_t) ent, (uint32_t) k);

            for (uint32_t i = lane; i < klen; i += W)
                huf[kst + i] = (uint16_t) kent;
            big &= big - 1;
        }
    }
    g_sync(g);
    *log_out = log;
    return used;
}

/* Huffman literals of one block -> dst (global).  The table lives in the idle ring region. */
template <int W>
CRYO_DEV int zsg_huffman_literals(ZswState &z, int lit_type, const uint8_t *p, uint32_t left,
                                  uint32_t regen, uint32_t streams, uint8_t *dst, uint8_t *smem,
                                  const Grp<W> &g)
{
    const uint32_t lane = g.lane;
    uint16_t *huf = reinterpret_cast<uint16_t *>(smem + ZSW_OFF_HUF);
    uint8_t  *work = smem + ZSW_OFF_HUFWORK;

    if (lit_type == 2)
    {
        int32_t  log = 0;
        uint32_t used = zsg_huf_build<W>(p, left, huf, work, &log, g);

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

        if (z.huf_log < 1 || zsg_huf_build<W>(z.huf_desc, z.huf_desc_len, huf, work, &log, g) == 0)
            return ST_FORMAT;
    }
    g_sync(g);
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
    return g_any(g, !ok) ? ST_FORMAT : ST_OK;
}

/* touch the lines of [p, p + n) so that later dependent reads find them in L2 */
template <int W>
CRYO_DEV void zsg_prefetch(const uint8_t *p, uint32_t n, const Grp<W> &g)
{
#ifndef CRYO_EMU
    for (uint32_t a = 128u * g.lane; a < n; a += 128u * W)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p + a));
#else
    (void) p;
    (void) n;
    (void) g;
#endif
}

template <int W>
CRYO_DEV void zsg_lits_fill(ZswLits &L, uint32_t ip, const Grp<W> &g)
{
    g_sync(g);
    L.wbase = ip & ~15u;
    L.wvalid = true;
    /* all loads first, then the stores: one memory round trip per refill */
    constexpr uint32_t R = ZSW_LITWIN / 16u / W;
    uint4 v[R];

#pragma unroll
    for (uint32_t k = 0; k < R; k++)
    {
        const uint32_t a = L.wbase + 16u * (g.lane + k * W);

        v[k] = a < L.lim ? ld16(L.abase + a) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (uint32_t k = 0; k < R; k++)
        st16(L.win + 16u * (g.lane + k * W), v[k]);
    g_sync(g);
}

template <int W>
CRYO_DEV void zsg_lits_emit(GOut &o, ZswLits &L, uint32_t n, const Grp<W> &g)
{
    const uint32_t lane = g.lane;
    if (n == 0)
        return;
    if (L.rle)
        gx_fill_byte(o, L.rle_byte, n, g);
    else if (n >= GX_BULK || n + 16u > ZSW_LITWIN)
    {
        /* long run: straight from global memory (through the ring when it is short of a bulk) */
        if (n >= GX_BULK)
            gx_literals(o, L.abase + L.delta + L.pos, n, g);
        else
        {
            const uint8_t *gp = L.abase + L.delta + L.pos;

            for (uint32_t i = lane; i < n; i += W)
                o.ring[(o.pos + i) & GX_RMASK] = gp[i];
            o.pos += n;
            g_sync(g);
            gx_drain(o, g);
        }
    }
    else
    {
        uint32_t ip = L.delta + L.pos;

        if (!L.wvalid || ip + n > L.wbase + ZSW_LITWIN)
            zsg_lits_fill<W>(L, ip, g);
        gx_literals(o, L.win + (ip - L.wbase), n, g);
    }
    L.pos += n;
}

template <int W>
CRYO_DEV void zsg_bits_fill(ZswBits &B, const Grp<W> &g)
{
    const uint32_t topbyte = B.bitpos >> 3;     /* highest byte still needed */

    g_sync(g);
    B.wlo = topbyte + 16u > ZSW_SEQWIN ? ((topbyte + 16u - ZSW_SEQWIN) & ~15u) : 0u;
    {
        constexpr uint32_t R = ZSW_SEQWIN / 16u / W;
        uint4 v[R];

#pragma unroll
        for (uint32_t k = 0; k < R; k++)
        {
            const uint32_t a = B.wlo + 16u * (g.lane + k * W);

            v[k] = a < B.lim ? ld16(B.abase + a) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (uint32_t k = 0; k < R; k++)
            st16(B.win + 16u + 16u * (g.lane + k * W), v[k]);
    }
    g_sync(g);
    if (B.wlo == 0)
        for (uint32_t i = g.lane; i < (B.lowbit >> 3); i += W)
            B.win[16u + i] = 0;                 /* bytes in front of the stream read as zero */
    g_sync(g);
}

template <int W>
CRYO_DEV bool zsg_bits_init(ZswBits &B, const uint8_t *p, uint32_t n, uint8_t *win, const Grp<W> &g)
{
    const uint32_t lane = g.lane;
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
    if (g.lane < 4)
        reinterpret_cast<uint32_t *>(win)[g.lane] = 0;
    zsg_bits_fill<W>(B, g);
    return true;
}

/*
 * Decode the zstd frame(s) at src[0, csize) into out[0, cap).  One warp; `smem` is this
 * warp's ZSW_PER_WARP bytes; `scratch` is ZSTDD_SCRATCH_BYTES of global memory private to
 * this warp (16-byte aligned); predef holds the three predefined FSE tables.
 */
template <int W>
CRYO_DEV void zstdg_decode_frame(const uint8_t *src, uint32_t csize, uint8_t *out, uint32_t cap,
                                 uint32_t *out_size, int32_t *status, uint8_t *scratch,
                                 const uint32_t *predef, uint8_t *smem, const Grp<W> &g)
{
    const uint32_t lane = g.lane;
    GOut     o;
    ZswState z;
    int      err = ST_OK;
    uint32_t ip = 0;
    const uint8_t *in = src;
    uint32_t *ll_tab = reinterpret_cast<uint32_t *>(smem + ZSW_OFF_LL);
    uint32_t *of_tab = reinterpret_cast<uint32_t *>(smem + ZSW_OFF_OF);
    uint32_t *ml_tab = reinterpret_cast<uint32_t *>(smem + ZSW_OFF_ML);

    gx_init(o, out, cap, smem + ZSW_OFF_RING);
    zsg_prefetch<W>(src, csize, g);

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
                    gx_literals(o, in + ip, bsize, g);
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
                    gx_fill_byte(o, in[ip], bsize, g);
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
                    gx_drain_all(o, g);          /* the Huffman table overlays the ring */
                if (lt >= 2)
                {
                    err = zsg_huffman_literals<W>(z, (int) lt, bp + lhdr, lcsize, regen, streams,
                                               scratch, smem, g);
                    if (err != ST_OK)
                        break;
                }
                if (nseq)
                {
                    uint8_t *work = smem + ZSW_OFF_FSEWORK;
                    uint32_t u;

                    g_sync(g);
                    u = zsg_seq_table<W>(z, (modes >> 6) & 3, 0, bp + sp, bsize - sp, ll_tab, work, predef,
                                      z.ll_log, g);
                    if (u == ~0u)
                    {
                        err = ST_FORMAT;
                        break;
                    }
                    sp += u;
                    u = zsg_seq_table<W>(z, (modes >> 4) & 3, 1, bp + sp, bsize - sp, of_tab, work, predef,
                                      z.of_log, g);
                    if (u == ~0u)
                    {
                        err = ST_FORMAT;
                        break;
                    }
                    sp += u;
                    u = zsg_seq_table<W>(z, (modes >> 2) & 3, 2, bp + sp, bsize - sp, ml_tab, work, predef,
                                      z.ml_log, g);
                    if (u == ~0u)
                    {
                        err = ST_FORMAT;
                        break;
                    }
                    sp += u;
                }
                if (lt >= 2)
                    gx_after_bulk(o, 0, g);      /* the ring is garbage now: re-prime it */

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

                    if (sp > bsize || !zsg_bits_init<W>(B, bp + sp, bsize - sp, smem + ZSW_OFF_SEQWIN, g))
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
                     * W sequences at a time.  Pass 1 (group-uniform, serial): walk the three FSE
                     * states; only the state bits are read here, lane k keeps the cells and the
                     * bit position of sequence k.  Pass 2 (one sequence per lane): every lane
                     * extracts its own offset / match-length / literal-length extra bits.  Then
                     * the W sequences are executed in order (repeat offsets, checks, copies).
                     */
                    for (uint32_t done = 0; done < nseq && err == ST_OK; done += W)
                    {
                        const uint32_t gn = nseq - done < (uint32_t) W ? nseq - done : (uint32_t) W;
                        uint32_t my_cl = 0, my_co = 0, my_cm = 0, my_bp = 0;
                        int32_t  under = 0;

                        /* W sequences take at most W x 89 bits; keep them and a peek inside the window */
                        if (B.wlo != 0 && B.bitpos < B.wlo * 8u + 96u * W + 128u)
                            zsg_bits_fill<W>(B, g);
                        const int32_t wbits = (int32_t) (B.wlo * 8u);

                        for (uint32_t k = 0; k < gn; k++)
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

                        if (lane < gn)
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
                        for (uint32_t k = 0; k < gn; k++)
                        {
                            const uint32_t ov = g_shfl(g, my_ov, k);
                            const uint32_t ml = g_shfl(g, my_ml, k);
                            const uint32_t ll = g_shfl(g, my_ll, k);
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
                             * in the ring: byte moves inside shared memory, W per step.
                             */
                            const uint32_t lip = L.delta + lpos;

                            if (ll <= 64u && ml <= 64u && !L.rle && off >= ml && off <= GX_RING - 64u &&
                                mpos - off >= o.lo)
                            {
                                if (ll)
                                {
                                    if (!L.wvalid || lip + ll > L.wbase + ZSW_LITWIN || lip < L.wbase)
                                        zsg_lits_fill<W>(L, lip, g);
                                    for (uint32_t i = lane; i < ll; i += W)
                                        o.ring[(o.pos + i) & GX_RMASK] = L.win[lip - L.wbase + i];
                                    g_sync(g);
                                }
                                for (uint32_t i = lane; i < ml; i += W)
                                    o.ring[(mpos + i) & GX_RMASK] = o.ring[(mpos - off + i) & GX_RMASK];
                                o.pos = epos;
                                lpos += ll;
                                g_sync(g);
                                if (o.pos - o.flushed >= (16u * W))
                                    gx_drain(o, g);
                                continue;
                            }
                            L.pos = lpos;
                            zsg_lits_emit<W>(o, L, ll, g);
                            lpos += ll;
                            gx_match(o, off, ml, g);
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
                zsg_lits_emit<W>(o, L, rest, g);
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
    gx_drain_all(o, g);
    if (lane == 0)
    {
        *out_size = err == ST_OK ? o.pos : 0u;
        *status = err;
    }
}

