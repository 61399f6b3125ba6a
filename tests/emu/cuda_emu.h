/*
 * tests/emu/cuda_emu.h -- a small SIMT emulator for running the .cuh kernel bodies
 * on the CPU inside the `-m "not gpu"` test-suite (TEST INFRASTRUCTURE ONLY).
 *
 * There is no GPU in the build container, so kernel logic (barrier protocol,
 * warp shuffles/ballots, shared-memory indexing, bounds) is exercised here
 * before it goes to a B200.  Every CUDA thread of a CTA is a ucontext fiber on
 * ONE OS thread; __syncthreads/__syncwarp/__shfl_sync/__ballot_sync are
 * cooperative barriers between fibers.  CTAs run one after another.  The
 * emulator detects barrier deadlocks (a barrier no fiber can complete).
 *
 * This is not a product path: nothing in pg_cryogen_b200/ or libcryogpu.so
 * includes it; libcryogpu.so has no CPU fallback.
 */
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct uint3e { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
static inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return uint4{a, b, c, d}; }
static inline uint2 make_uint2(uint32_t a, uint32_t b) { return uint2{a, b}; }

namespace emu {

struct Fiber
{
    ucontext_t  ctx;
    char       *stack;
    bool        done;
    unsigned    tid;
};

/* one barrier instance per participation mask, so that sub-warp groups (e.g. four 8-lane
 * groups using 0xFF << 8g) synchronise independently, as they do on the hardware */
struct MaskBarrier
{
    unsigned    mask = 0;
    int         gen = 0;
    int         arrived = 0;
};

struct WarpState
{
    std::vector<MaskBarrier> bars;
    uint64_t    slot[32];
};

struct Cta
{
    std::vector<Fiber> fibers;
    std::vector<WarpState> warps;
    int         nthreads = 0;
    int         alive = 0;
    int         bar_gen = 0, bar_arrived = 0;
    int         named_gen[16] = {0}, named_arrived[16] = {0};
    uint8_t    *smem = nullptr;
    unsigned    cur = 0;
    dim3        block_idx, block_dim, grid_dim;
    ucontext_t  sched;
    long        idle_switches = 0;
    std::function<void()> body;
};

static Cta *g_cta = nullptr;

static inline void
yield()
{
    Cta *c = g_cta;
    swapcontext(&c->fibers[c->cur].ctx, &c->sched);
}

static inline void
progress()
{
    g_cta->idle_switches = 0;
}

static void
fiber_entry()
{
    Cta *c = g_cta;

    c->body();
    c->fibers[c->cur].done = true;
    c->alive--;
    progress();
    swapcontext(&c->fibers[c->cur].ctx, &c->sched);
}

static inline void
run_cta(Cta &c)
{
    const size_t STK = 256 * 1024;

    g_cta = &c;
    c.fibers.resize(c.nthreads);
    c.warps.assign((c.nthreads + 31) / 32, WarpState());
    c.alive = c.nthreads;
    for (int t = 0; t < c.nthreads; t++)
    {
        Fiber &f = c.fibers[t];

        f.stack = (char *) malloc(STK);
        f.done = false;
        f.tid = t;
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = STK;
        f.ctx.uc_link = nullptr;
        makecontext(&f.ctx, fiber_entry, 0);
    }
    c.idle_switches = 0;
    while (c.alive > 0)
    {
        for (int t = 0; t < c.nthreads && c.alive > 0; t++)
        {
            if (c.fibers[t].done)
                continue;
            c.cur = t;
            c.idle_switches++;
            swapcontext(&c.sched, &c.fibers[t].ctx);
            if (c.idle_switches > 64L * c.nthreads + 1024)
            {
                fprintf(stderr, "cuda_emu: DEADLOCK in block %u (no fiber can make progress)\n",
                        c.block_idx.x);
                abort();
            }
        }
    }
    for (auto &f : c.fibers)
        free(f.stack);
    g_cta = nullptr;
}

/* launch: body is invoked once per CUDA thread */
static inline void
launch(dim3 grid, dim3 block, size_t smem_bytes, std::function<void()> body)
{
    for (unsigned b = 0; b < grid.x; b++)
    {
        Cta c;

        c.nthreads = (int) block.x;
        c.block_idx = dim3(b, 0, 0);
        c.block_dim = block;
        c.grid_dim = grid;
        c.body = body;
        /* 16 guard bytes each side catch small overruns under ASan-less builds */
        uint8_t *raw = (uint8_t *) aligned_alloc(128, ((smem_bytes + 127) / 128) * 128 + 256);
        memset(raw, 0xCD, ((smem_bytes + 127) / 128) * 128 + 256);
        c.smem = raw + 128;
        run_cta(c);
        for (int i = 0; i < 128; i++)
            if (raw[i] != 0xCD || raw[128 + ((smem_bytes + 127) / 128) * 128 + i] != 0xCD)
            {
                fprintf(stderr, "cuda_emu: shared memory guard overwritten (block %u)\n", b);
                abort();
            }
        free(raw);
    }
}

static inline void
cta_barrier(int &gen, int &arrived, int count)
{
    int g = gen;

    progress();
    if (++arrived >= count)
    {
        arrived = 0;
        gen++;
        return;
    }
    while (gen == g)
        yield();
}

static inline void
warp_barrier(unsigned mask = 0xffffffffu)
{
    Cta *c = g_cta;
    WarpState &w = c->warps[c->cur / 32];
    int lanes = c->nthreads - (int) (c->cur / 32) * 32;

    if (lanes < 32)
        mask &= (1u << lanes) - 1u;
    if (!(mask & (1u << (c->cur % 32))))
    {
        fprintf(stderr, "cuda_emu: lane %u calls a warp collective without being in its mask %08x\n",
                c->cur % 32, mask);
        abort();
    }
    size_t k = 0;

    for (; k < w.bars.size(); k++)
        if (w.bars[k].mask == mask)
            break;
    if (k == w.bars.size())
    {
        w.bars.emplace_back();
        w.bars[k].mask = mask;
    }
    /* index, not reference: other fibers may grow the vector while this one waits */
    int g = w.bars[k].gen;

    progress();
    if (++w.bars[k].arrived >= __builtin_popcount(mask))
    {
        w.bars[k].arrived = 0;
        w.bars[k].gen++;
        return;
    }
    while (c->warps[c->cur / 32].bars[k].gen == g)
        yield();
}

struct Tidx { unsigned x, y, z; };
static inline Tidx tidx() { return Tidx{g_cta->cur, 0, 0}; }

} /* namespace emu */

#define threadIdx (emu::tidx())
#define blockIdx (emu::g_cta->block_idx)
#define blockDim (emu::g_cta->block_dim)
#define gridDim (emu::g_cta->grid_dim)

static inline void __syncthreads() { emu::cta_barrier(emu::g_cta->bar_gen, emu::g_cta->bar_arrived, emu::g_cta->nthreads); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::warp_barrier(mask); }
/* __syncthreads_or: barrier + OR of the predicate over the CTA (two barriers: collect, then read and reset) */
static inline int emu_syncthreads_or(int pred)
{
    static int acc = 0, result = 0;

    if (pred)
        acc = 1;
    emu::cta_barrier(emu::g_cta->bar_gen, emu::g_cta->bar_arrived, emu::g_cta->nthreads);
    if (emu::g_cta->cur == 0)
    {
        result = acc;
        acc = 0;
    }
    emu::cta_barrier(emu::g_cta->bar_gen, emu::g_cta->bar_arrived, emu::g_cta->nthreads);
    return result;
}
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __nanosleep(unsigned) { emu::yield(); }

/* bar.sync id, count -- named barrier over `count` threads */
static inline void emu_named_barrier(int id, int count)
{
    emu::cta_barrier(emu::g_cta->named_gen[id], emu::g_cta->named_arrived[id], count);
}

template <typename T>
static inline T emu_exchange(unsigned mask, T v, int src_lane)
{
    emu::Cta *c = emu::g_cta;
    uint64_t raw = 0;

    memcpy(&raw, &v, sizeof(T));
    c->warps[c->cur / 32].slot[c->cur % 32] = raw;
    emu::warp_barrier(mask);
    raw = c->warps[c->cur / 32].slot[src_lane & 31];
    emu::warp_barrier(mask);
    T r;
    memcpy(&r, &raw, sizeof(T));
    return r;
}

/* width: the warp is cut into segments of `width` lanes; src / delta are relative to the segment */
template <typename T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{
    int lane = emu::g_cta->cur % 32, base = lane & ~(width - 1);
    return emu_exchange(mask, v, base + (src & (width - 1)));
}
template <typename T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32)
{
    int lane = emu::g_cta->cur % 32, base = lane & ~(width - 1);
    return emu_exchange(mask, v, lane - (int) d >= base ? lane - (int) d : lane);
}
template <typename T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32)
{
    int lane = emu::g_cta->cur % 32, base = lane & ~(width - 1);
    return emu_exchange(mask, v, lane + (int) d < base + width ? lane + (int) d : lane);
}
template <typename T> static inline T __shfl_xor_sync(unsigned mask, T v, int m, int width = 32)
{
    int lane = emu::g_cta->cur % 32;
    (void) width;
    return emu_exchange(mask, v, lane ^ m);
}
static inline unsigned __ballot_sync(unsigned mask, int pred)
{
    emu::Cta *c = emu::g_cta;
    int lanes = c->nthreads - (int) (c->cur / 32) * 32;
    unsigned r = 0;

    if (lanes > 32)
        lanes = 32;
    c->warps[c->cur / 32].slot[c->cur % 32] = pred ? 1 : 0;
    emu::warp_barrier(mask);
    for (int i = 0; i < lanes; i++)
        if (mask & (1u << i))
            r |= (unsigned) c->warps[c->cur / 32].slot[i] << i;
    emu::warp_barrier(mask);
    return r;
}
static inline unsigned __match_any_sync(unsigned mask, unsigned v)
{
    emu::Cta *c = emu::g_cta;
    unsigned r = 0;

    c->warps[c->cur / 32].slot[c->cur % 32] = v;
    emu::warp_barrier(mask);
    for (int i = 0; i < 32; i++)
        if ((mask & (1u << i)) && (unsigned) c->warps[c->cur / 32].slot[i] == v)
            r |= 1u << i;
    emu::warp_barrier(mask);
    return r;
}
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
static inline int __all_sync(unsigned m, int p) { return __ballot_sync(m, !p) == 0; }
template <typename F> static inline unsigned emu_reduce(unsigned mask, unsigned v, F f)
{
    emu::Cta *c = emu::g_cta;
    unsigned r = 0;
    bool first = true;

    c->warps[c->cur / 32].slot[c->cur % 32] = v;
    emu::warp_barrier(mask);
    for (int i = 0; i < 32; i++)
        if (mask & (1u << i))
        {
            unsigned x = (unsigned) c->warps[c->cur / 32].slot[i];
            r = first ? x : f(r, x);
            first = false;
        }
    emu::warp_barrier(mask);
    return r;
}
static inline unsigned __reduce_add_sync(unsigned m, unsigned v) { return emu_reduce(m, v, [](unsigned a, unsigned b) { return a + b; }); }
static inline unsigned __reduce_max_sync(unsigned m, unsigned v) { return emu_reduce(m, v, [](unsigned a, unsigned b) { return a > b ? a : b; }); }
static inline unsigned __reduce_min_sync(unsigned m, unsigned v) { return emu_reduce(m, v, [](unsigned a, unsigned b) { return a < b ? a : b; }); }
static inline unsigned __reduce_or_sync(unsigned m, unsigned v) { return emu_reduce(m, v, [](unsigned a, unsigned b) { return a | b; }); }

static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned) v) : 32; }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long) v) : 64; }
static inline unsigned __brev(unsigned v)
{
    unsigned r = 0;
    for (int i = 0; i < 32; i++)
        r |= ((v >> i) & 1u) << (31 - i);
    return r;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s)
{
    uint64_t v = ((uint64_t) hi << 32) | lo;
    return (unsigned) (v >> (s & 31));
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s)
{
    uint64_t v = ((uint64_t) hi << 32) | lo;
    return (unsigned) ((v << (s & 31)) >> 32);
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel)
{
    uint64_t v = ((uint64_t) b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++)
        r |= (unsigned) ((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned) (((uint64_t) a * b) >> 32); }

template <typename T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> static inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> static inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> static inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
template <typename T> static inline T atomicCAS(T *p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }
template <typename T> static inline T __ldg(const T *p) { return *p; }
template <typename T> static inline T __ldcs(const T *p) { return *p; }
template <typename T> static inline void __stcs(T *p, T v) { *p = v; }
template <typename T> static inline T min(T a, T b) { return a < b ? a : b; }
template <typename T> static inline T max(T a, T b) { return a > b ? a : b; }
