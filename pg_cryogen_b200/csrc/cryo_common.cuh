/*
 * cryo_common.cuh -- device-side building blocks shared by the sm_100a codec kernels:
 * 16-byte global/shared access, byte-shifted vector loads, and "team" primitives
 * (copy / fill executed cooperatively by a group of threads with coalesced
 * 16-byte stores).
 *
 * The same source is compiled by nvcc for sm_100a (the product) and, with
 * -DCRYO_EMU, by g++ on top of tests/emu/cuda_emu.h for the CPU test-suite.
 * The emulated build is test scaffolding only; libcryogpu.so has no CPU path.
 */
#pragma once
#include <stdint.h>

#ifdef CRYO_EMU
#include "cuda_emu.h"
#define CRYO_DEV static inline
#define CRYO_SMEM_BASE() (emu::g_cta->smem)
#else
#include <cuda_runtime.h>
#define CRYO_DEV __device__ __forceinline__
extern __shared__ __align__(128) uint8_t cryo_dyn_smem[];
#define CRYO_SMEM_BASE() (cryo_dyn_smem)
#endif

#define CRYO_FULL 0xffffffffu

/* per-block status values: keep in sync with include/cryogpu.h */
#define ST_OK 0
#define ST_INPUT 1
#define ST_OUTPUT 2
#define ST_OFFSET 3
#define ST_FORMAT 4
#define ST_SIZE 5
#define ST_METHOD 6
#define ST_UNSUPPORTED 7

CRYO_DEV uint4 ld16(const uint8_t *p) { return *reinterpret_cast<const uint4 *>(p); }
CRYO_DEV void st16(uint8_t *p, uint4 v) { *reinterpret_cast<uint4 *>(p) = v; }
CRYO_DEV uint32_t ld4(const uint8_t *p) { return *reinterpret_cast<const uint32_t *>(p); }

CRYO_DEV uint32_t align_down16(uint32_t v) { return v & ~15u; }

/*
 * Bytes [sh, sh+16) of the 32-byte little-endian string A:B (0 < sh < 16 is the
 * useful range; sh == 0 returns A).  sh is warp-uniform in every caller, so the
 * switch does not diverge.
 */
CRYO_DEV uint4 shift_combine(uint4 A, uint4 B, uint32_t sh)
{
    uint32_t w0, w1, w2, w3, w4;
    uint32_t bs = (sh & 3u) * 8u;

    switch (sh >> 2)
    {
        case 0: w0 = A.x; w1 = A.y; w2 = A.z; w3 = A.w; w4 = B.x; break;
        case 1: w0 = A.y; w1 = A.z; w2 = A.w; w3 = B.x; w4 = B.y; break;
        case 2: w0 = A.z; w1 = A.w; w2 = B.x; w3 = B.y; w4 = B.z; break;
        default: w0 = A.w; w1 = B.x; w2 = B.y; w3 = B.z; w4 = B.w; break;
    }
    return make_uint4(__funnelshift_r(w0, w1, bs), __funnelshift_r(w1, w2, bs),
                      __funnelshift_r(w2, w3, bs), __funnelshift_r(w3, w4, bs));
}

/*
 * team_copy: n bytes src -> dst, any alignment, ranges must not overlap.
 * Executed by a team of nthr (>= 16) threads, this thread being tid.
 * Stores are 16-byte aligned vectors (plus <= 15 head and tail bytes); when
 * src and dst are not congruent mod 16 each vector is assembled from two
 * aligned loads with funnel shifts.  The aligned loads may touch up to 15
 * bytes either side of [src, src+n) inside the same 16-byte granules.
 */
CRYO_DEV void team_copy(uint8_t *dst, const uint8_t *src, uint32_t n, uint32_t tid, uint32_t nthr)
{
    if (n < 64)
    {
        for (uint32_t i = tid; i < n; i += nthr)
            dst[i] = src[i];
        return;
    }
    uint32_t head = (16u - (uint32_t) ((uintptr_t) dst & 15u)) & 15u;
    if (tid < head)
        dst[tid] = src[tid];
    uint32_t nvec = (n - head) >> 4;
    const uint8_t *s = src + head;
    uint8_t *d = dst + head;
    uint32_t sh = (uint32_t) ((uintptr_t) s & 15u);
    uint32_t v = tid;

    if (sh == 0)
    {
        for (; v + 3 * nthr < nvec; v += 4 * nthr)
        {
            uint4 a = ld16(s + 16 * (size_t) v);
            uint4 b = ld16(s + 16 * (size_t) (v + nthr));
            uint4 c = ld16(s + 16 * (size_t) (v + 2 * nthr));
            uint4 e = ld16(s + 16 * (size_t) (v + 3 * nthr));
            st16(d + 16 * (size_t) v, a);
            st16(d + 16 * (size_t) (v + nthr), b);
            st16(d + 16 * (size_t) (v + 2 * nthr), c);
            st16(d + 16 * (size_t) (v + 3 * nthr), e);
        }
        for (; v < nvec; v += nthr)
            st16(d + 16 * (size_t) v, ld16(s + 16 * (size_t) v));
    }
    else
    {
        const uint8_t *sb = s - sh;

        for (; v + nthr < nvec; v += 2 * nthr)
        {
            uint4 a0 = ld16(sb + 16 * (size_t) v);
            uint4 a1 = ld16(sb + 16 * (size_t) v + 16);
            uint4 b0 = ld16(sb + 16 * (size_t) (v + nthr));
            uint4 b1 = ld16(sb + 16 * (size_t) (v + nthr) + 16);
            st16(d + 16 * (size_t) v, shift_combine(a0, a1, sh));
            st16(d + 16 * (size_t) (v + nthr), shift_combine(b0, b1, sh));
        }
        for (; v < nvec; v += nthr)
        {
            uint4 a0 = ld16(sb + 16 * (size_t) v);
            uint4 a1 = ld16(sb + 16 * (size_t) v + 16);
            st16(d + 16 * (size_t) v, shift_combine(a0, a1, sh));
        }
    }
    uint32_t done = head + (nvec << 4);
    if (tid < n - done)
        dst[done + tid] = src[done + tid];
}

/* team_fill_byte: n bytes of value b at dst (any alignment). */
CRYO_DEV void team_fill_byte(uint8_t *dst, uint8_t b, uint32_t n, uint32_t tid, uint32_t nthr)
{
    if (n < 64)
    {
        for (uint32_t i = tid; i < n; i += nthr)
            dst[i] = b;
        return;
    }
    uint32_t head = (16u - (uint32_t) ((uintptr_t) dst & 15u)) & 15u;
    if (tid < head)
        dst[tid] = b;
    uint32_t nvec = (n - head) >> 4;
    uint8_t *d = dst + head;
    uint32_t w = b * 0x01010101u;
    uint4 val = make_uint4(w, w, w, w);

    if (n >= 8192u)
    {
        /* a long run streams through L2 (evict-first): what the same SMs are reading meanwhile
         * (literals, sequences, match sources) should stay there */
        for (uint32_t v = tid; v < nvec; v += nthr)
            __stcs(reinterpret_cast<uint4 *>(d + 16 * (size_t) v), val);
    }
    else
        for (uint32_t v = tid; v < nvec; v += nthr)
            st16(d + 16 * (size_t) v, val);
    uint32_t done = head + (nvec << 4);
    if (tid < n - done)
        dst[done + tid] = b;
}

/*
 * team_fill_from_pattern: dst[i] = pat[(phase + i) % plen] for i < n, where pat
 * is a shared-memory buffer holding plen pattern bytes followed by at least 19
 * bytes of wrap-around (pat[plen + j] == pat[j]).  plen >= 64.  dst any alignment.
 */
CRYO_DEV void team_fill_from_pattern(uint8_t *dst, const uint8_t *pat, uint32_t plen,
                                     uint32_t phase, uint32_t n, uint32_t tid, uint32_t nthr)
{
    uint32_t head = (16u - (uint32_t) ((uintptr_t) dst & 15u)) & 15u;

    if (head > n)
        head = n;
    if (tid < head)
        dst[tid] = pat[(phase + tid) % plen];
    uint32_t nvec = (n - head) >> 4;
    uint8_t *d = dst + head;
    uint32_t idx = (phase + head + 16u * tid) % plen;
    uint32_t stride = (16u * nthr) % plen;

    for (uint32_t v = tid; v < nvec; v += nthr)
    {
        const uint8_t *p = pat + (idx & ~3u);
        uint32_t bs = (idx & 3u) * 8u;
        uint32_t w0 = ld4(p), w1 = ld4(p + 4), w2 = ld4(p + 8), w3 = ld4(p + 12), w4 = ld4(p + 16);

        st16(d + 16 * (size_t) v,
             make_uint4(__funnelshift_r(w0, w1, bs), __funnelshift_r(w1, w2, bs),
                        __funnelshift_r(w2, w3, bs), __funnelshift_r(w3, w4, bs)));
        idx += stride;
        if (idx >= plen)
            idx -= plen;
    }
    uint32_t done = head + (nvec << 4);
    if (tid < n - done)
        dst[done + tid] = pat[(phase + done + tid) % plen];
}
