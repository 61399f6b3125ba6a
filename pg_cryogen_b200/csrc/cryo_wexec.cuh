/*
 * cryo_wexec.cuh -- warp-level output executor ("one warp owns one cryo block").
 *
 * The sequence stream of an LZ4 block / zstd block is inherently serial, so the unit
 * that walks it is a single warp.  What a B200 offers is many such warps at once:
 * this executor keeps the per-warp footprint small (a WX_RING-byte ring in shared
 * memory) so that many blocks are in flight per SM, and it has no CTA-wide barrier.
 *
 *   - short literal runs and short matches are appended to the ring, which mirrors
 *     the most recent output; match sources inside the ring are served from shared
 *     memory (tens of cycles) instead of an L2 round trip;
 *   - the ring drains to global memory in 512-byte pieces: one coalesced 16-byte
 *     store per lane;
 *   - anything long (the >1 MB zero run of a sparse block, long literal runs, RLE
 *     blocks) bypasses the ring: the warp streams aligned 16-byte vectors straight
 *     to global memory; overlapping matches become periodic pattern fills whose
 *     period is staged in the (then idle) ring.
 *
 * All functions are warp-collective: every lane calls them with identical
 * arguments (the state is warp-uniform and lives in registers).
 */
#pragma once
#include "cryo_common.cuh"

#ifndef WX_RING
#define WX_RING   2048u           /* power of two */
#endif
#define WX_RMASK  (WX_RING - 1u)
#define WX_DRAIN  512u            /* bytes per drain step (32 lanes x 16 B) */
#define WX_BULK   512u            /* runs at least this long bypass the ring */
#define WX_PAT_MAXOFF 512u       /* k*off + 32 must fit the ring */

struct WOut
{
    uint8_t    *out;            /* global output block, 16-byte aligned */
    uint8_t    *ring;           /* shared, WX_RING bytes, 16-byte aligned */
    uint32_t    cap;
    uint32_t    pos;            /* next output byte */
    uint32_t    flushed;        /* multiple of 16; out[0, flushed) is in global memory */
    uint32_t    lo;             /* ring holds out[max(lo, pos - WX_RING + 64), pos) */
};

CRYO_DEV void wx_init(WOut &o, uint8_t *out, uint32_t cap, uint8_t *ring)
{
    o.out = out;
    o.cap = cap;
    o.ring = ring;
    o.pos = 0;
    o.flushed = 0;
    o.lo = 0;
}

/* drain whole 512-byte pieces; callers have __syncwarp()'ed after their ring writes */
CRYO_DEV void wx_drain(WOut &o, uint32_t lane)
{
    if (o.pos - o.flushed < WX_DRAIN)
        return;
    do
    {
        uint32_t a = o.flushed + 16u * lane;

        st16(o.out + a, ld16(o.ring + (a & WX_RMASK)));
        o.flushed += WX_DRAIN;
    } while (o.pos - o.flushed >= WX_DRAIN);
}

/* drain everything up to pos (end of block, or before a bulk operation) */
CRYO_DEV void wx_drain_all(WOut &o, uint32_t lane)
{
    uint32_t p0 = o.pos & ~15u;

    __syncwarp();               /* ring writes of other lanes */
    for (uint32_t a = o.flushed + 16u * lane; a < p0; a += 512u)
        st16(o.out + a, ld16(o.ring + (a & WX_RMASK)));
    if (lane < o.pos - p0)
        o.out[p0 + lane] = o.ring[(p0 + lane) & WX_RMASK];
    o.flushed = p0;
    __syncwarp();
}

/* after n bytes were written at pos directly to global memory */
CRYO_DEV void wx_after_bulk(WOut &o, uint32_t n, uint32_t lane)
{
    o.pos += n;
    o.flushed = o.pos & ~15u;
    o.lo = o.flushed;
    __syncwarp();
    if (lane < o.pos - o.flushed)
        o.ring[(o.flushed + lane) & WX_RMASK] = o.out[o.flushed + lane];
    __syncwarp();
}

/* the same when the bytes written were all b: the ring's tail needs no read-back (a load of bytes stored a moment
 * ago waits for the store to land: the zero runs of a sparse block each cost a round trip that way) */
CRYO_DEV void wx_after_fill(WOut &o, uint32_t n, uint8_t b, uint32_t lane)
{
    o.pos += n;
    o.flushed = o.pos & ~15u;
    __syncwarp();
    if (n >= WX_RING)
    {
        /* the ring's whole reach lies inside the run: it can serve the matches that copy from the run's end
         * (the first tuples after the zero run of a sparse block do) without a read of global memory */
        const uint32_t w = (uint32_t) b * 0x01010101u;

        for (uint32_t k = 16u * lane; k < WX_RING; k += 512u)
            st16(o.ring + k, make_uint4(w, w, w, w));
        o.lo = o.pos - (WX_RING - 64u);
    }
    else
    {
        o.lo = o.flushed;
        if (lane < o.pos - o.flushed)
            o.ring[(o.flushed + lane) & WX_RMASK] = b;
    }
    __syncwarp();
}

/* n literal bytes from src (shared or global memory, readable by every lane) */
CRYO_DEV void wx_literals(WOut &o, const uint8_t *src, uint32_t n, uint32_t lane)
{
    if (n >= WX_BULK)
    {
        wx_drain_all(o, lane);
        team_copy(o.out + o.pos, src, n, lane, 32);
        wx_after_bulk(o, n, lane);
        return;
    }
    for (uint32_t i = lane; i < n; i += 32)
        o.ring[(o.pos + i) & WX_RMASK] = src[i];
    o.pos += n;
    __syncwarp();
    wx_drain(o, lane);
}

CRYO_DEV void wx_fill_byte(WOut &o, uint8_t b, uint32_t n, uint32_t lane)
{
    if (n >= 64)
    {
        wx_drain_all(o, lane);
        team_fill_byte(o.out + o.pos, b, n, lane, 32);
        wx_after_fill(o, n, b, lane);
        return;
    }
    for (uint32_t i = lane; i < n; i += 32)
        o.ring[(o.pos + i) & WX_RMASK] = b;
    o.pos += n;
    __syncwarp();
    wx_drain(o, lane);
}

/* long match on global memory; out[0, pos) is in global memory and visible */
CRYO_DEV void wx_bulk_match(WOut &o, uint32_t off, uint32_t n, uint32_t lane)
{
    uint8_t *dst = o.out + o.pos;

    if (off >= n)
    {
        team_copy(dst, dst - off, n, lane, 32);
        return;
    }
    if (off == 1)
    {
        team_fill_byte(dst, dst[-1], n, lane, 32);
        return;
    }
    if (off <= 16 && (16 % off) == 0 && n >= 64)
    {
        /* the period divides 16: every aligned 16-byte vector of the run is the same */
        uint32_t head = (16u - (uint32_t) ((uintptr_t) dst & 15u)) & 15u;
        const uint8_t *src = dst - off;
        uint32_t w[4];

#pragma unroll
        for (uint32_t q = 0; q < 4; q++)
        {
            uint32_t v = 0;

#pragma unroll
            for (uint32_t j = 0; j < 4; j++)
                v |= (uint32_t) src[(head + 4 * q + j) % off] << (8 * j);
            w[q] = v;
        }
        if (lane < head)
            dst[lane] = src[lane % off];
        uint32_t nvec = (n - head) >> 4;
        uint8_t *d = dst + head;
        uint4 val = make_uint4(w[0], w[1], w[2], w[3]);

        for (uint32_t v = lane; v < nvec; v += 32)
            st16(d + 16 * (size_t) v, val);
        uint32_t done = head + (nvec << 4);

        if (lane < n - done)
            dst[done + lane] = src[(done + lane) % off];
        return;
    }
    if (off < WX_PAT_MAXOFF)
    {
        /* stage k whole periods (k*off >= WX_PAT_MAXOFF) in the idle ring */
        uint32_t k = (WX_PAT_MAXOFF + off - 1) / off;
        uint32_t plen = k * off;
        const uint8_t *src = dst - off;
        uint32_t r = lane % off;

        for (uint32_t j = lane; j < plen + 32; j += 32)
        {
            o.ring[j] = src[r];
            r += 32 % off;
            r = r >= off ? r - off : r;
        }
        __syncwarp();
        team_fill_from_pattern(dst, o.ring, plen, 0, n, lane, 32);
        __syncwarp();
        return;
    }
    /* long period: every round copies the largest whole number of periods available */
    uint32_t done = 0;

    while (done < n)
    {
        uint32_t avail = ((off + done) / off) * off;
        uint32_t m = n - done < avail ? n - done : avail;

        team_copy(dst + done, dst + done - avail, m, lane, 32);
        done += m;
        __syncwarp();
    }
}

/* match: out[pos+i] = out[pos+i-off], i < n; caller validated off and the bounds */
CRYO_DEV void wx_match(WOut &o, uint32_t off, uint32_t n, uint32_t lane)
{
    if (n >= WX_BULK)
    {
        if (off == 1 && o.pos > o.lo)
        {
            /* a run of the previous byte, which the ring still holds */
            const uint8_t b = o.ring[(o.pos - 1u) & WX_RMASK];

            wx_drain_all(o, lane);
#ifndef ZP_ABLATE_FILLS         /* timing experiment only: the long runs are not written */
            team_fill_byte(o.out + o.pos, b, n, lane, 32);
#endif
            wx_after_fill(o, n, b, lane);
            return;
        }
        wx_drain_all(o, lane);
        wx_bulk_match(o, off, n, lane);
        wx_after_bulk(o, n, lane);
        return;
    }
    const uint32_t src = o.pos - off;
    const bool in_ring = off <= WX_RING - 64u;
    /*
     * The steps of a copy (32 bytes each) must not overtake one another when a later step's destination falls on the
     * ring slot an earlier step still has to read: the match overlaps itself (off < n), or it is so far back that
     * the destination wraps round onto its own source (off + n beyond the ring).  The second case went unsynchronised
     * until the end of round 2: lockstep execution hid it on the device, the emulator, which runs the lanes one after
     * the other between barriers, did not (a 490-byte match at distance 1 878).
     */
    const bool ordered = off < n || off + n + 64u > WX_RING;

    if (in_ring && src >= o.lo)
    {
        /* the whole source is in the ring: shared memory only */
        if (off >= 32 || off >= n)
        {
            for (uint32_t i0 = 0; i0 < n; i0 += 32)      /* uniform trip count */
            {
                uint32_t i = i0 + lane;

                if (i < n)
                    o.ring[(o.pos + i) & WX_RMASK] = o.ring[(src + i) & WX_RMASK];
                if (ordered)
                    __syncwarp();
            }
        }
        else
        {
            uint32_t r = lane % off;
            const uint32_t step = 32 % off;

            for (uint32_t i = lane; i < n; i += 32)
            {
                o.ring[(o.pos + i) & WX_RMASK] = o.ring[(src + r) & WX_RMASK];
                r += step;
                r = r >= off ? r - off : r;
            }
        }
        o.pos += n;
        __syncwarp();
        wx_drain(o, lane);
        return;
    }
    if (off >= n && (!in_ring || src + n <= o.lo))
    {
        /* the whole source is in global memory and does not overlap the destination: issue
         * every load before the first store, so a run costs one memory round trip */
        for (uint32_t i0 = 0; i0 < n; i0 += 128)
        {
            uint8_t b[4];

#pragma unroll
            for (uint32_t k = 0; k < 4; k++)
            {
                const uint32_t i = i0 + 32 * k + lane;

                b[k] = i < n ? o.out[src + i] : (uint8_t) 0;
            }
#pragma unroll
            for (uint32_t k = 0; k < 4; k++)
            {
                const uint32_t i = i0 + 32 * k + lane;

                if (i < n)
                    o.ring[(o.pos + i) & WX_RMASK] = b[k];
            }
        }
    }
    else if (off >= 32 || off >= n)
    {
        for (uint32_t i0 = 0; i0 < n; i0 += 32)
        {
            uint32_t i = i0 + lane;

            if (i < n)
            {
                uint32_t s = src + i;
                uint8_t  b = (in_ring && s >= o.lo) ? o.ring[s & WX_RMASK] : o.out[s];

                o.ring[(o.pos + i) & WX_RMASK] = b;
            }
            if (ordered)
                __syncwarp();
        }
    }
    else
    {
        /* short period: every byte comes from the off bytes before pos */
        uint32_t r = lane % off;
        const uint32_t step = 32 % off;

        for (uint32_t i = lane; i < n; i += 32)
        {
            uint32_t s = src + r;
            uint8_t  b = s >= o.lo ? o.ring[s & WX_RMASK] : o.out[s];

            o.ring[(o.pos + i) & WX_RMASK] = b;
            r += step;
            r = r >= off ? r - off : r;
        }
    }
    o.pos += n;
    __syncwarp();
    wx_drain(o, lane);
}
