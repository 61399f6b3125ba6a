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
#include "zstd_decode.cuh"

#define ZSW_WARPS     2
#define ZSW_THREADS   (32 * ZSW_WARPS)
#define ZSW_LITWIN    1024u
/* per-warp shared memory: [0,6K) ring+litwin / Huffman table+scratch, then the FSE tables */
#define ZSW_OFF_RING    0
#define ZSW_OFF_LITWIN  WX_RING                 /* sequence phase */
#define ZSW_OFF_HUF     0                       /* entropy phase: u16[2048] */
#define ZSW_OFF_WORK    4096                    /* entropy phase: 2 KB */
#define ZSW_OFF_LL      6144
#define ZSW_OFF_OF      (ZSW_OFF_LL + 2048)
#define ZSW_OFF_ML      (ZSW_OFF_OF + 1024)
#define ZSW_PER_WARP    (ZSW_OFF_ML + 2048)     /* 11264 */
#define ZSW_SMEM        (ZSW_WARPS * ZSW_PER_WARP)
#define ZSW_PREDEF_CELLS (64 + 32 + 64)         /* LL, OF, ML predefined tables */

#if WX_RING + ZSW_LITWIN > 6144
#error "ring + literal window must fit the 6 KB they share with the Huffman table"
#endif

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
};

/* one sequence table (t: 0 LL, 1 OF, 2 ML); returns bytes of description consumed or ~0u */
CRYO_DEV uint32_t zsw_seq_table(int mode, int t, const uint8_t *p, uint32_t left, uint32_t *cell,
                                uint8_t *work, const uint32_t *predef, int &logv, uint32_t lane)
{
    const int max_log = t == 1 ? 8 : 9, max_sym = t == 0 ? 35 : t == 1 ? 31 : 52;
    int16_t  *counts = reinterpret_cast<int16_t *>(work + ZW_COUNTS);
    uint16_t *next = reinterpret_cast<uint16_t *>(work + ZW_NEXT);
    uint16_t *cum = reinterpret_cast<uint16_t *>(work + ZW_NEXT + 128);

    switch (mode)
    {
        case 0:
        {
            const uint32_t n = t == 1 ? 32u : 64u, o = t == 0 ? 0u : t == 1 ? 64u : 96u;

            for (uint32_t i = lane; i < n; i += 32)
                cell[i] = predef[o + i];
            logv = t == 1 ? 5 : 6;
            __syncwarp();
            return 0;
        }
        case 1:
            if (left < 1 || p[0] > max_sym)
                return ~0u;
            if (lane == 0)
                cell[0] = p[0];                       /* nbits 0, base 0 */
            logv = 0;
            __syncwarp();
            return 1;
        case 2:
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
            return used;
        }
        default:
            return logv < 0 ? ~0u : 0u;
    }
}

/* Huffman literals of one block -> dst (global).  Table + scratch live in the idle ring. */
CRYO_DEV int zsw_huffman_literals(ZswState &z, int lit_type, const uint8_t *p, uint32_t left,
                                  uint32_t regen, uint32_t streams, uint8_t *dst, uint8_t *smem,
                                  uint32_t lane)
{
    uint16_t *huf = reinterpret_cast<uint16_t *>(smem + ZSW_OFF_HUF);
    uint8_t  *work = smem + ZSW_OFF_WORK;

    if (lit_type == 2)
    {
        int32_t  log = 0;
        uint32_t used = huf_build_table(p, left, huf, work, &log, lane);

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

        if (z.huf_log < 1 || huf_build_table(z.huf_desc, z.huf_desc_len, huf, work, &log, lane) == 0)
            return ST_FORMAT;
    }
    __syncwarp();
    bool ok = true;

    if (streams == 1)
    {
        if (lane == 0)
            ok = huf_decode_stream(huf, z.huf_log, p, left, dst, regen);
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

                ok = huf_decode_stream(huf, z.huf_log, p + 6 + so, sn, dst + lane * seg, cnt);
            }
        }
    }
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

CRYO_DEV void zsw_lits_emit(WOut &o, ZswLits &L, uint32_t n, uint32_t lane)
{
    if (n == 0)
        return;
    if (L.rle)
        wx_fill_byte(o, L.rle_byte, n, lane);
    else if (n >= WX_BULK)
        wx_literals(o, L.abase + L.delta + L.pos, n, lane);
    else
    {
        uint32_t ip = L.delta + L.pos;

        if (!L.wvalid || ip + n > L.wbase + ZSW_LITWIN)
        {
            __syncwarp();
            L.wbase = ip & ~15u;
            L.wvalid = true;
#pragma unroll
            for (uint32_t k = 0; k < ZSW_LITWIN / 512; k++)
            {
                uint32_t a = L.wbase + 512u * k + 16u * lane;

                if (a < L.lim)
                    st16(L.win + 512u * k + 16u * lane, ld16(L.abase + a));
            }
            __syncwarp();
        }
        wx_literals(o, L.win + (ip - L.wbase), n, lane);
    }
    L.pos += n;
}

/*
 * Decode the zstd frame(s) at src[0, csize) into out[0, cap).  One warp; `smem` is this
 * warp's ZSW_PER_WARP bytes; `scratch` is ZSTDD_SCRATCH_BYTES of global memory private
 * to this warp (16-byte aligned); predef holds the three predefined FSE tables.
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
                {
                    wx_drain_all(o, lane);          /* the Huffman table overlays the ring */
                    err = zsw_huffman_literals(z, (int) lt, bp + lhdr, lcsize, regen, streams,
                                               scratch, smem, lane);
                    if (err != ST_OK)
                        break;
                }
                if (nseq)
                {
                    uint8_t *work = smem + ZSW_OFF_WORK;
                    uint32_t u;

                    u = zsw_seq_table((modes >> 6) & 3, 0, bp + sp, bsize - sp, ll_tab, work, predef,
                                      z.ll_log, lane);
                    if (u == ~0u)
                    {
                        err = ST_FORMAT;
                        break;
                    }
                    sp += u;
                    u = zsw_seq_table((modes >> 4) & 3, 1, bp + sp, bsize - sp, of_tab, work, predef,
                                      z.of_log, lane);
                    if (u == ~0u)
                    {
                        err = ST_FORMAT;
                        break;
                    }
                    sp += u;
                    u = zsw_seq_table((modes >> 2) & 3, 2, bp + sp, bsize - sp, ml_tab, work, predef,
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
                    BitsBack bb;
                    const int ll_log = z.ll_log, of_log = z.of_log, ml_log = z.ml_log;

                    if (sp > bsize || !bb_init(bb, bp + sp, bsize - sp))
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
                            z.rep2 = z.rep1;
                            z.rep1 = z.rep0;
                            z.rep0 = off;
                        }
                        else
                        {
                            uint32_t idx = ov - 1 + (ll == 0 ? 1u : 0u);

                            if (idx == 0)
                                off = z.rep0;
                            else
                            {
                                off = idx == 1 ? z.rep1 : idx == 2 ? z.rep2 : z.rep0 - 1;
                                if (idx > 1)
                                    z.rep2 = z.rep1;
                                z.rep1 = z.rep0;
                                z.rep0 = off;
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
                        if ((uint64_t) o.pos + ll + ml > cap)
                        {
                            err = ST_OUTPUT;
                            break;
                        }
                        if (o.pos + ll + ml - block_start > block_max)
                        {
                            err = ST_FORMAT;
                            break;
                        }
                        zsw_lits_emit(o, L, ll, lane);
                        if (off == 0 || off > o.pos - frame_start)
                        {
                            err = ST_OFFSET;
                            break;
                        }
                        wx_match(o, off, ml, lane);
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
