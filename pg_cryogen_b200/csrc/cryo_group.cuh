/*
 * cryo_group.cuh -- sub-warp groups: W consecutive lanes (W = 8 or 32) that own one cryo block.
 *
 * The walk over an LZ4 / zstd sequence stream is serial, and on one warp per block every
 * lane repeats the same scalar work: the kernels were bound by instruction issue, not by
 * memory (profiles/ r01d: 50 % issue-slot use at 32 % of HBM).  With W = 8 a warp carries
 * four blocks at once -- the scalar walk of four blocks costs one warp instruction instead
 * of four -- while the shared-memory footprint per block, which bounds the blocks in flight
 * per SM, stays the same.  All collectives below take the group's participation mask, so the
 * four groups of a warp synchronise independently (independent thread scheduling) and may
 * diverge freely; blocks of one table have the same shape, so they mostly do not.
 */
#pragma once
#include "cryo_common.cuh"

template <int W>
struct Grp
{
    uint32_t    lane;           /* 0 .. W-1 */
    uint32_t    base;           /* first lane of the group inside its warp */
    uint32_t    mask;           /* participation mask inside the warp */
};

template <int W>
CRYO_DEV Grp<W> grp_make(uint32_t warp_lane)
{
    Grp<W> g;

    g.lane = warp_lane & (uint32_t) (W - 1);
    g.base = warp_lane & ~(uint32_t) (W - 1);
    g.mask = W == 32 ? 0xffffffffu : (((1u << (W & 31)) - 1u) << g.base);
    return g;
}

template <int W> CRYO_DEV void g_sync(const Grp<W> &g) { __syncwarp(g.mask); }

template <int W> CRYO_DEV uint32_t g_ballot(const Grp<W> &g, bool p)
{
    return W == 32 ? __ballot_sync(g.mask, p) : ((__ballot_sync(g.mask, p) >> g.base) & ((1u << (W & 31)) - 1u));
}
template <int W> CRYO_DEV bool g_any(const Grp<W> &g, bool p) { return __ballot_sync(g.mask, p) != 0; }
template <int W> CRYO_DEV bool g_all(const Grp<W> &g, bool p) { return __ballot_sync(g.mask, !p) == 0; }

template <int W> CRYO_DEV uint32_t g_shfl(const Grp<W> &g, uint32_t v, uint32_t k)
{
    return __shfl_sync(g.mask, v, (int) k, W);
}
template <int W> CRYO_DEV uint32_t g_shfl_up(const Grp<W> &g, uint32_t v, uint32_t d)
{
    return __shfl_up_sync(g.mask, v, d, W);
}
template <int W> CRYO_DEV uint32_t g_reduce_add(const Grp<W> &g, uint32_t v) { return __reduce_add_sync(g.mask, v); }
template <int W> CRYO_DEV uint32_t g_reduce_or(const Grp<W> &g, uint32_t v) { return __reduce_or_sync(g.mask, v); }
template <int W> CRYO_DEV uint32_t g_reduce_max(const Grp<W> &g, uint32_t v) { return __reduce_max_sync(g.mask, v); }

/* lanes of the group (group-relative bits) that hold the same value */
template <int W> CRYO_DEV uint32_t g_match_any(const Grp<W> &g, uint32_t v)
{
    const uint32_t m = __match_any_sync(g.mask, v);

    return W == 32 ? m : ((m >> g.base) & ((1u << (W & 31)) - 1u));
}

/* ---- team primitives for any team size (cryo_common.cuh's need >= 16 threads) ------------ */

/* n bytes src -> dst, any alignment, no overlap; 16-byte aligned stores in the body */
template <int W>
CRYO_DEV void g_copy(uint8_t *dst, const uint8_t *src, uint32_t n, uint32_t tid)
{
    if (n < 64)
    {
        for (uint32_t i = tid; i < n; i += W)
            dst[i] = src[i];
        return;
    }
    const uint32_t head = (16u - (uint32_t) ((uintptr_t) dst & 15u)) & 15u;

    for (uint32_t i = tid; i < head; i += W)
        dst[i] = src[i];
    const uint32_t nvec = (n - head) >> 4;
    const uint8_t *s = src + head;
    uint8_t       *d = dst + head;
    const uint32_t sh = (uint32_t) ((uintptr_t) s & 15u);

    if (sh == 0)
    {
        uint32_t v = tid;

        for (; v + 3 * W < nvec; v += 4 * W)
        {
            uint4 a = ld16(s + 16 * (size_t) v);
            uint4 b = ld16(s + 16 * (size_t) (v + W));
            uint4 c = ld16(s + 16 * (size_t) (v + 2 * W));
            uint4 e = ld16(s + 16 * (size_t) (v + 3 * W));

            st16(d + 16 * (size_t) v, a);
            st16(d + 16 * (size_t) (v + W), b);
            st16(d + 16 * (size_t) (v + 2 * W), c);
            st16(d + 16 * (size_t) (v + 3 * W), e);
        }
        for (; v < nvec; v += W)
            st16(d + 16 * (size_t) v, ld16(s + 16 * (size_t) v));
    }
    else
    {
        const uint8_t *sb = s - sh;

        for (uint32_t v = tid; v < nvec; v += W)
        {
            uint4 a0 = ld16(sb + 16 * (size_t) v);
            uint4 a1 = ld16(sb + 16 * (size_t) v + 16);

            st16(d + 16 * (size_t) v, shift_combine(a0, a1, sh));
        }
    }
    const uint32_t done = head + (nvec << 4);

    for (uint32_t i = done + tid; i < n; i += W)
        dst[i] = src[i];
}

/* n bytes of value b at dst (any alignment) */
template <int W>
CRYO_DEV void g_fill_byte(uint8_t *dst, uint8_t b, uint32_t n, uint32_t tid)
{
    if (n < 64)
    {
        for (uint32_t i = tid; i < n; i += W)
            dst[i] = b;
        return;
    }
    const uint32_t head = (16u - (uint32_t) ((uintptr_t) dst & 15u)) & 15u;

    for (uint32_t i = tid; i < head; i += W)
        dst[i] = b;
    const uint32_t nvec = (n - head) >> 4;
    uint8_t       *d = dst + head;
    const uint32_t w = b * 0x01010101u;
    const uint4    val = make_uint4(w, w, w, w);
    uint32_t       v = tid;

    for (; v + 3 * W < nvec; v += 4 * W)
    {
        st16(d + 16 * (size_t) v, val);
        st16(d + 16 * (size_t) (v + W), val);
        st16(d + 16 * (size_t) (v + 2 * W), val);
        st16(d + 16 * (size_t) (v + 3 * W), val);
    }
    for (; v < nvec; v += W)
        st16(d + 16 * (size_t) v, val);
    const uint32_t done = head + (nvec << 4);

    for (uint32_t i = done + tid; i < n; i += W)
        dst[i] = b;
}

/*
 * dst[i] = pat[(phase + i) % plen] for i < n; pat is a shared-memory buffer holding plen
 * pattern bytes followed by at least 19 bytes of wrap-around.  plen >= 64.
 */
template <int W>
CRYO_DEV void g_fill_from_pattern(uint8_t *dst, const uint8_t *pat, uint32_t plen, uint32_t phase,
                                  uint32_t n, uint32_t tid)
{
    uint32_t head = (16u - (uint32_t) ((uintptr_t) dst & 15u)) & 15u;

    if (head > n)
        head = n;
    for (uint32_t i = tid; i < head; i += W)
        dst[i] = pat[(phase + i) % plen];
    const uint32_t nvec = (n - head) >> 4;
    uint8_t       *d = dst + head;
    uint32_t       idx = (phase + head + 16u * tid) % plen;
    const uint32_t stride = (16u * W) % plen;

    for (uint32_t v = tid; v < nvec; v += W)
    {
        const uint8_t *p = pat + (idx & ~3u);
        const uint32_t bs = (idx & 3u) * 8u;
        const uint32_t w0 = ld4(p), w1 = ld4(p + 4), w2 = ld4(p + 8), w3 = ld4(p + 12), w4 = ld4(p + 16);

        st16(d + 16 * (size_t) v,
             make_uint4(__funnelshift_r(w0, w1, bs), __funnelshift_r(w1, w2, bs),
                        __funnelshift_r(w2, w3, bs), __funnelshift_r(w3, w4, bs)));
        idx += stride;
        if (idx >= plen)
            idx -= plen;
    }
    const uint32_t done = head + (nvec << 4);

    for (uint32_t i = done + tid; i < n; i += W)
        dst[i] = pat[(phase + i) % plen];
}
