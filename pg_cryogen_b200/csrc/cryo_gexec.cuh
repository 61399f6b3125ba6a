/*
 * cryo_gexec.cuh -- output executor for a group of W lanes that owns one cryo block
 * (cryo_group.cuh).  Same design as cryo_wexec.cuh, with the group width a template
 * parameter: a GX_RING-byte ring in shared memory mirrors the most recent output, short
 * literal runs and matches are byte moves inside it, it drains to global memory in pieces of
 * W x 16 bytes (one 16-byte store per lane), and anything long bypasses it -- aligned 16-byte
 * stores straight to global memory, overlapping matches as periodic pattern fills.
 *
 * Every function is group-collective: all W lanes call it with identical arguments.
 */
#pragma once
#include "cryo_group.cuh"

#define GX_RING   2048u
#define GX_RMASK  (GX_RING - 1u)
#define GX_BULK   512u            /* runs at least this long bypass the ring */
#define GX_PAT_MAXOFF 512u       /* k*off + 32 must fit the ring */

struct GOut
{
    uint8_t    *out;            /* global output block, 16-byte aligned */
    uint8_t    *ring;           /* shared, GX_RING bytes, 16-byte aligned */
    uint32_t    cap;
    uint32_t    pos;            /* next output byte */
    uint32_t    flushed;        /* multiple of 16; out[0, flushed) is in global memory */
    uint32_t    lo;             /* ring holds out[max(lo, pos - GX_RING + 64), pos) */
};

CRYO_DEV void gx_init(GOut &o, uint8_t *out, uint32_t cap, uint8_t *ring)
{
    o.out = out;
    o.cap = cap;
    o.ring = ring;
    o.pos = 0;
    o.flushed = 0;
    o.lo = 0;
}

/* drain whole pieces of W x 16 bytes; callers have synchronised after their ring writes */
template <int W>
CRYO_DEV void gx_drain(GOut &o, const Grp<W> &g)
{
    while (o.pos - o.flushed >= 16u * W)
    {
        const uint32_t a = o.flushed + 16u * g.lane;

        st16(o.out + a, ld16(o.ring + (a & GX_RMASK)));
        o.flushed += 16u * W;
    }
}

/* drain everything up to pos (end of block, or before a bulk operation) */
template <int W>
CRYO_DEV void gx_drain_all(GOut &o, const Grp<W> &g)
{
    const uint32_t p0 = o.pos & ~15u;

    g_sync(g);
    for (uint32_t a = o.flushed + 16u * g.lane; a < p0; a += 16u * W)
        st16(o.out + a, ld16(o.ring + (a & GX_RMASK)));
    for (uint32_t i = p0 + g.lane; i < o.pos; i += W)
        o.out[i] = o.ring[i & GX_RMASK];
    o.flushed = p0;
    g_sync(g);
}

/* after n bytes were written at pos directly to global memory */
template <int W>
CRYO_DEV void gx_after_bulk(GOut &o, uint32_t n, const Grp<W> &g)
{
    o.pos += n;
    o.flushed = o.pos & ~15u;
    o.lo = o.flushed;
    g_sync(g);
    for (uint32_t i = o.flushed + g.lane; i < o.pos; i += W)
        o.ring[i & GX_RMASK] = o.out[i];
    g_sync(g);
}

/* n literal bytes from src (shared or global memory, readable by every lane) */
template <int W>
CRYO_DEV void gx_literals(GOut &o, const uint8_t *src, uint32_t n, const Grp<W> &g)
{
    if (n >= GX_BULK)
    {
        gx_drain_all(o, g);
        g_copy<W>(o.out + o.pos, src, n, g.lane);
        gx_after_bulk(o, n, g);
        return;
    }
    for (uint32_t i = g.lane; i < n; i += W)
        o.ring[(o.pos + i) & GX_RMASK] = src[i];
    o.pos += n;
    g_sync(g);
    gx_drain(o, g);
}

template <int W>
CRYO_DEV void gx_fill_byte(GOut &o, uint8_t b, uint32_t n, const Grp<W> &g)
{
    if (n >= 64)
    {
        gx_drain_all(o, g);
        g_fill_byte<W>(o.out + o.pos, b, n, g.lane);
        gx_after_bulk(o, n, g);
        return;
    }
    for (uint32_t i = g.lane; i < n; i += W)
        o.ring[(o.pos + i) & GX_RMASK] = b;
    o.pos += n;
    g_sync(g);
    gx_drain(o, g);
}

/* long match on global memory; out[0, pos) is in global memory and visible */
template <int W>
CRYO_DEV void gx_bulk_match(GOut &o, uint32_t off, uint32_t n, const Grp<W> &g)
{
    uint8_t *dst = o.out + o.pos;
    const uint32_t lane = g.lane;

    if (off >= n)
    {
        g_copy<W>(dst, dst - off, n, lane);
        return;
    }
    if (off == 1)
    {
        g_fill_byte<W>(dst, dst[-1], n, lane);
        return;
    }
    if (off <= 16 && (16 % off) == 0 && n >= 64)
    {
        /* the period divides 16: every aligned 16-byte vector of the run is the same */
        const uint32_t head = (16u - (uint32_t) ((uintptr_t) dst & 15u)) & 15u;
        const uint8_t *src = dst - off;
        uint32_t       w[4];

#pragma unroll
        for (uint32_t q = 0; q < 4; q++)
        {
            uint32_t v = 0;

#pragma unroll
            for (uint32_t j = 0; j < 4; j++)
                v |= (uint32_t) src[(head + 4 * q + j) % off] << (8 * j);
            w[q] = v;
        }
        for (uint32_t i = lane; i < head; i += W)
            dst[i] = src[i % off];
        const uint32_t nvec = (n - head) >> 4;
        uint8_t       *d = dst + head;
        const uint4    val = make_uint4(w[0], w[1], w[2], w[3]);

        for (uint32_t v = lane; v < nvec; v += W)
            st16(d + 16 * (size_t) v, val);
        const uint32_t done = head + (nvec << 4);

        for (uint32_t i = done + lane; i < n; i += W)
            dst[i] = src[i % off];
        return;
    }
    if (off < GX_PAT_MAXOFF)
    {
        /* stage k whole periods (k*off >= GX_PAT_MAXOFF) in the idle ring */
        const uint32_t k = (GX_PAT_MAXOFF + off - 1) / off;
        const uint32_t plen = k * off;
        const uint8_t *src = dst - off;

        for (uint32_t j = lane; j < plen + 32; j += W)
            o.ring[j] = src[j % off];
        g_sync(g);
        g_fill_from_pattern<W>(dst, o.ring, plen, 0, n, lane);
        g_sync(g);
        return;
    }
    /* long period: every round copies the largest whole number of periods available */
    uint32_t done = 0;

    while (done < n)
    {
        const uint32_t avail = ((off + done) / off) * off;
        const uint32_t m = n - done < avail ? n - done : avail;

        g_copy<W>(dst + done, dst + done - avail, m, lane);
        done += m;
        __threadfence_block();
        g_sync(g);
    }
}

/* match: out[pos+i] = out[pos+i-off], i < n; caller validated off and the bounds */
template <int W>
CRYO_DEV void gx_match(GOut &o, uint32_t off, uint32_t n, const Grp<W> &g)
{
    const uint32_t lane = g.lane;

    if (n >= GX_BULK)
    {
        gx_drain_all(o, g);
        __threadfence_block();
        gx_bulk_match(o, off, n, g);
        gx_after_bulk(o, n, g);
        return;
    }
    const uint32_t src = o.pos - off;
    const bool     in_ring = off <= GX_RING - 64u;

    if (off >= n)
    {
        /* source and destination do not overlap: all lanes move at once */
        if (in_ring && src >= o.lo)
            for (uint32_t i = lane; i < n; i += W)
                o.ring[(o.pos + i) & GX_RMASK] = o.ring[(src + i) & GX_RMASK];
        else if (src + n <= o.lo || !in_ring)
        {
            /* the whole source is in global memory: issue every load before the first store so
             * that a run costs one memory round trip, not one per W bytes */
            for (uint32_t i0 = 0; i0 < n; i0 += 8 * W)
            {
                uint8_t b[8];

#pragma unroll
                for (uint32_t k = 0; k < 8; k++)
                {
                    const uint32_t i = i0 + k * W + lane;

                    b[k] = i < n ? o.out[src + i] : (uint8_t) 0;
                }
#pragma unroll
                for (uint32_t k = 0; k < 8; k++)
                {
                    const uint32_t i = i0 + k * W + lane;

                    if (i < n)
                        o.ring[(o.pos + i) & GX_RMASK] = b[k];
                }
            }
        }
        else
            for (uint32_t i = lane; i < n; i += W)
            {
                const uint32_t s = src + i;

                o.ring[(o.pos + i) & GX_RMASK] = (in_ring && s >= o.lo) ? o.ring[s & GX_RMASK] : o.out[s];
            }
    }
    else
    {
        /* overlap: every byte comes from the off bytes before pos (period off) */
        uint32_t r = lane % off;
        const uint32_t step = W % off;

        for (uint32_t i = lane; i < n; i += W)
        {
            const uint32_t s = src + r;

            o.ring[(o.pos + i) & GX_RMASK] = (in_ring && s >= o.lo) ? o.ring[s & GX_RMASK] : o.out[s];
            r += step;
            r = r >= off ? r - off : r;
        }
    }
    o.pos += n;
    g_sync(g);
    gx_drain(o, g);
}
