/*
 * cryo_cx.cuh -- CTA-cooperative output executor ("one CTA owns one cryo block").
 *
 * The warp executor (cryo_wexec.cuh) walks the sequences of a block one after the other:
 * fine when thousands of blocks are in flight, hopeless for the call the reference actually
 * makes (one block per call, cache.c:178) and for match-rich blocks (120 K sequences at one
 * L2 round trip each).  Here the whole CTA works on ONE block and the sequences themselves are
 * the unit of parallelism:
 *
 *   - the front end (lz4_decode_c.cuh, zstd_decode_c.cuh) hands over CX_THREADS sequences at a
 *     time, one per thread: literal length, match length, offset, where the literals are;
 *   - output positions are a CTA-wide prefix sum, so every sequence knows where it writes
 *     before any byte moves;
 *   - literals depend on nothing: all threads place theirs at once (a lane for a short run,
 *     the warp for a medium one);
 *   - a match byte is a copy of an earlier byte.  If that byte was produced before the chunk it is
 *     final and is copied at once.  If it lies inside the chunk, the byte only records how far
 *     back its source is (a 16-bit distance per byte of the chunk, dl[]), and the chunk is then
 *     resolved by pointer jumping: in a round every unresolved byte looks at its source -- final:
 *     take the value; unresolved: add the source's distance to its own -- so a chain of n copies
 *     resolves in log2(n) rounds of perfectly regular work, whatever the structure of the block
 *     (a dense low-cardinality cryo block has chains hundreds of matches long: executing them
 *     match by match, in dependency order, was 8 x slower, profiles/r02_cx_phases.txt);
 *   - the most recent output lives in a CX_RING-byte ring in shared memory, so a match source
 *     is a shared-memory read (LZ4's whole 64 KiB window fits); the ring drains to HBM as
 *     aligned 16-byte vectors by all threads;
 *   - runs of CX_BIG bytes or more (the ~1 MB zero run of a sparse block, RLE blocks, long
 *     literal runs) bypass the ring: the whole CTA streams them to global memory, overlapping
 *     matches as periodic fills, and the ring's history is reloaded afterwards.
 *
 * All functions are CTA-collective: every thread calls them, uniform arguments are marked.
 */
#pragma once
#include "cryo_common.cuh"

#ifndef CX_THREADS
#define CX_THREADS 1024u
#endif
#define CX_WARPS   (CX_THREADS / 32u)
#define CX_RING    131072u              /* power of two; >= CX_HIST + CX_SPAN + slack */
#define CX_RMASK   (CX_RING - 1u)
#define CX_SPAN    16384u               /* output bytes one chunk of sequences may cover (dl[]: 2 bytes each) */
#define CX_BIG     8192u                /* runs at least this long are CTA-wide bulk operations (2 * CX_BIG <= CX_SPAN) */
#define CX_LANE    32u                  /* runs up to this long are moved by the sequence's own lane */
#define CX_HIST    65536u               /* history reloaded into the ring after a bulk operation (LZ4: offsets <= 65535) */
#define CX_FAR     49152u               /* a match from further back is copied when it is met (CX_FAR + CX_BIG < 65536, CX_FAR >= CX_SPAN) */
#define CX_PAT     1024u                /* pattern staging for periodic fills (k * off >= 512, + 32) */
#define CX_TPW     ((CX_SPAN / 512u + CX_WARPS - 1u) / CX_WARPS)    /* sixteen 32-byte rows of a chunk per warp, times this (1 with 1 024 threads) */
#define CX_ROWSTEP (32u * CX_WARPS)     /* bytes between two rows of one warp */

#ifdef CRYO_EMU
#define CX_LDCG(p) (*(p))
#define CX_VLD(p) (*reinterpret_cast<const volatile uint32_t *>(p))
#define CX_VST(p, v) (*reinterpret_cast<volatile uint32_t *>(p) = (v))
#define CX_SYNC_OR(p) emu_syncthreads_or(p)
#define CX_COMPILER_FENCE() __asm__ __volatile__("" ::: "memory")
#else
#define CX_LDCG(p) __ldcg(p)
#define CX_VLD(p) (*reinterpret_cast<const volatile uint32_t *>(p))
#define CX_VST(p, v) (*reinterpret_cast<volatile uint32_t *>(p) = (v))
#define CX_SYNC_OR(p) __syncthreads_or(p)
#define CX_COMPILER_FENCE() asm volatile("" ::: "memory")
#endif

/* development aid (-DCX_PROF): thread 0 of CTA 0 charges the cycles since the previous mark to a phase counter */
#if defined(CX_PROF) && !defined(CRYO_EMU)
__device__ unsigned long long cx_prof[32];
__device__ long long cx_prof_last;
#define CXP(i)                                                     \
    if (threadIdx.x == 0 && blockIdx.x == 0)                       \
    {                                                              \
        const long long t_ = clock64();                            \
        cx_prof[i] += (unsigned long long) (t_ - cx_prof_last);    \
        cx_prof_last = t_;                                         \
    }
#define CXP_COUNT(i, v)                                            \
    if (threadIdx.x == 0 && blockIdx.x == 0)                       \
        cx_prof[i] += (v);
#else
#define CXP(i)
#define CXP_COUNT(i, v)
#endif

/*
 * A run of more than CX_LANE bytes is moved by a whole warp.  Left to the warp that holds the sequence, a chunk
 * of few long sequences (a literal-heavy block: two dozen sequences of 600 literal bytes per chunk, all of them
 * in warp 0) kept one warp busy and thirty-one waiting: 1.1 M of the 3.9 M cycles of a dense hex block.  Such runs
 * are posted as jobs and taken by all warps in turn; what does not fit the list stays with the owning warp.
 */
#ifndef CX_JOBS
#define CX_JOBS    384u
#endif
struct CxJob
{
    uint32_t    pos, len;               /* output position, bytes */
    uint32_t    x;                      /* literals: fill byte or ~0u (copy from p); match: offset */
    uint32_t    kind;                   /* 0: literals, 1: match */
    const uint8_t *p;                   /* literals to copy */
};

/* shared-memory control block of one CTA */
struct CxSh
{
    uint32_t wsa[CX_WARPS + 1], wsb[CX_WARPS + 1];     /* scan scratch: per-warp sums, [CX_WARPS] = total */
    uint32_t cut;                       /* first sequence of the chunk that is not executed in it */
    uint32_t err;                       /* (sequence index << 8 | status) of the first failing sequence, ~0u: none */
    uint32_t bcast[4];
    uint32_t njobs;                     /* jobs posted for the chunk (may exceed CX_JOBS: those stay with their warp) */
    CxJob    jobs[CX_JOBS];
};

/* per-thread view; every field is uniform over the CTA */
struct Cx
{
    uint8_t    *out;                    /* global output block, 16-byte aligned */
    uint8_t    *ring;                   /* shared, CX_RING bytes, 16-byte aligned */
    uint8_t    *pat;                    /* shared, CX_PAT + 32 bytes, 16-byte aligned */
    uint16_t   *dl;                     /* shared, CX_SPAN entries, 16-byte aligned: distance to the source of every
                                         * unresolved byte of the chunk, 0: the byte is final */
    CxSh       *sh;
    uint32_t    cap;
    uint32_t    pos;                    /* next output byte */
    uint32_t    flushed;                /* multiple of 16; out[0, flushed) is in global memory */
    uint32_t    ring_lo;                /* the ring holds out[max(ring_lo, newest - CX_RING), pos) */
    int         err;
};

CRYO_DEV void cx_init(Cx &cx, uint8_t *out, uint32_t cap, uint8_t *ring, uint8_t *pat, uint16_t *dl, CxSh *sh)
{
    cx.out = out;
    cx.cap = cap;
    cx.ring = ring;
    cx.pat = pat;
    cx.dl = dl;
    cx.sh = sh;
    cx.pos = 0;
    cx.flushed = 0;
    cx.ring_lo = 0;
    cx.err = ST_OK;
}

/*
 * Inclusive prefix sums of two values over the CTA (thread order).  Totals are left in
 * sh->wsa[CX_WARPS] / sh->wsb[CX_WARPS].  Two barriers; the caller keeps a barrier between two
 * calls (every use below has one).
 */
CRYO_DEV void cx_scan2(CxSh *sh, uint32_t &a, uint32_t &b, uint32_t tid)
{
    const uint32_t lane = tid & 31u, warp = tid >> 5;

#pragma unroll
    for (uint32_t d = 1; d < 32; d <<= 1)
    {
        const uint32_t x = __shfl_up_sync(CRYO_FULL, a, d), y = __shfl_up_sync(CRYO_FULL, b, d);

        if (lane >= d)
        {
            a += x;
            b += y;
        }
    }
    if (lane == 31)
    {
        sh->wsa[warp] = a;
        sh->wsb[warp] = b;
    }
    __syncthreads();
    if (warp == 0)
    {
        uint32_t x = lane < CX_WARPS ? sh->wsa[lane] : 0u, y = lane < CX_WARPS ? sh->wsb[lane] : 0u;
        const uint32_t x0 = x, y0 = y;

#pragma unroll
        for (uint32_t d = 1; d < 32; d <<= 1)
        {
            const uint32_t p = __shfl_up_sync(CRYO_FULL, x, d), q = __shfl_up_sync(CRYO_FULL, y, d);

            if (lane >= d)
            {
                x += p;
                y += q;
            }
        }
        if (lane < CX_WARPS)
        {
            sh->wsa[lane] = x - x0;     /* exclusive */
            sh->wsb[lane] = y - y0;
        }
        if (lane == CX_WARPS - 1u)
        {
            sh->wsa[CX_WARPS] = x;
            sh->wsb[CX_WARPS] = y;
        }
    }
    __syncthreads();
    a += sh->wsa[warp];
    b += sh->wsb[warp];
}

/* one byte of already produced output: the ring when it is there, global memory otherwise */
CRYO_DEV uint8_t cx_src(const Cx &cx, uint32_t x, uint32_t lo)
{
    return x >= lo ? cx.ring[x & CX_RMASK] : CX_LDCG(cx.out + x);
}

/* team copy of n bytes from src (global or shared) into the ring at output position p (split at the wrap) */
CRYO_DEV void cx_ring_put(Cx &cx, uint32_t p, const uint8_t *src, uint32_t n, uint32_t tid, uint32_t nthr)
{
    const uint32_t a = p & CX_RMASK, first = CX_RING - a < n ? CX_RING - a : n;

    team_copy(cx.ring + a, src, first, tid, nthr);
    if (first < n)
        team_copy(cx.ring, src + first, n - first, tid, nthr);
}

CRYO_DEV void cx_ring_fill(Cx &cx, uint32_t p, uint8_t b, uint32_t n, uint32_t tid, uint32_t nthr)
{
    const uint32_t a = p & CX_RMASK, first = CX_RING - a < n ? CX_RING - a : n;

    team_fill_byte(cx.ring + a, b, first, tid, nthr);
    if (first < n)
        team_fill_byte(cx.ring, b, n - first, tid, nthr);
}

/* drain out[flushed, pos & ~15) from the ring, all threads; ring writes are behind a barrier */
CRYO_DEV void cx_drain(Cx &cx, uint32_t tid)
{
    const uint32_t p0 = cx.pos & ~15u;

    for (uint32_t a = cx.flushed + 16u * tid; a < p0; a += 16u * CX_THREADS)
        st16(cx.out + a, ld16(cx.ring + (a & CX_RMASK)));
    cx.flushed = p0;
}

/* the (< 16) bytes between flushed and pos, bytewise (end of block, before a bulk operation) */
CRYO_DEV void cx_drain_tail(Cx &cx, uint32_t tid)
{
    if (tid < cx.pos - cx.flushed)
        cx.out[cx.flushed + tid] = cx.ring[(cx.flushed + tid) & CX_RMASK];
}

/* after a bulk operation wrote out[.., pos) straight to global memory: reload the ring's history */
CRYO_DEV void cx_reload(Cx &cx, uint32_t tid)
{
    const uint32_t lo = cx.pos > CX_HIST ? cx.pos - CX_HIST : 0u, a0 = lo & ~15u;

    __syncthreads();                    /* the bulk stores of every thread */
    for (uint32_t a = a0 + 16u * tid; a < cx.pos; a += 16u * CX_THREADS)
    {
        const uint4 v = CX_LDCG(reinterpret_cast<const uint4 *>(cx.out + a));

        st16(cx.ring + (a & CX_RMASK), v);      /* may run up to 15 bytes past pos: those slots are rewritten later */
    }
    cx.flushed = cx.pos & ~15u;
    cx.ring_lo = lo;
    __syncthreads();
}

/* long match straight on global memory by the whole CTA; out[0, pos) is in global memory */
CRYO_DEV void cx_bulk_match(Cx &cx, uint32_t off, uint32_t n, uint32_t tid)
{
    uint8_t *dst = cx.out + cx.pos;

    if (off >= n)
    {
        team_copy(dst, dst - off, n, tid, CX_THREADS);
        return;
    }
    if (off == 1)
    {
        team_fill_byte(dst, CX_LDCG(dst - 1), n, tid, CX_THREADS);
        return;
    }
    if (off < 512u)
    {
        /* k whole periods (k * off >= 512) staged in shared memory, then a pattern fill */
        const uint32_t k = (512u + off - 1u) / off, plen = k * off;
        const uint8_t *src = dst - off;

        for (uint32_t j = tid; j < plen + 32u; j += CX_THREADS)
            cx.pat[j] = CX_LDCG(src + j % off);
        __syncthreads();
        team_fill_from_pattern(dst, cx.pat, plen, 0, n, tid, CX_THREADS);
        return;
    }
    /* long period: every round copies the largest whole number of periods available */
    uint32_t done = 0;

    while (done < n)
    {
        const uint32_t avail = ((off + done) / off) * off;
        const uint32_t m = n - done < avail ? n - done : avail;

        team_copy(dst + done, dst + done - avail, m, tid, CX_THREADS);
        done += m;
        __syncthreads();
    }
}

/*
 * One sequence as a bulk operation (uniform arguments): ll literal bytes from lit (rle_byte >= 0:
 * ll copies of that byte), then a match of ml bytes at distance off.  Bounds were checked.
 */
CRYO_DEV void cx_bulk(Cx &cx, uint32_t ll, const uint8_t *lit, int rle_byte, uint32_t ml, uint32_t off, uint32_t tid)
{
    __syncthreads();                    /* ring writes of the previous chunk */
    cx_drain(cx, tid);
    cx_drain_tail(cx, tid);
    if (ll)
    {
        if (rle_byte >= 0)
            team_fill_byte(cx.out + cx.pos, (uint8_t) rle_byte, ll, tid, CX_THREADS);
        else
            team_copy(cx.out + cx.pos, lit, ll, tid, CX_THREADS);
        cx.pos += ll;
    }
    if (ml)
    {
        __syncthreads();                /* the match may read the literals and the drained tail */
        if (off == 1u && ml >= CX_HIST)
        {
            /* a run of one byte that covers the whole history window (the zero run of a sparse block): the ring's
             * history is that byte, no need to read it back from global memory */
            const uint8_t  b = CX_LDCG(cx.out + cx.pos - 1u);
            const uint32_t w = (uint32_t) b * 0x01010101u;

            team_fill_byte(cx.out + cx.pos, b, ml, tid, CX_THREADS);
            cx.pos += ml;
            const uint32_t lo = cx.pos - CX_HIST, a0 = lo & ~15u;

            for (uint32_t a = a0 + 16u * tid; a < cx.pos; a += 16u * CX_THREADS)
                st16(cx.ring + (a & CX_RMASK), make_uint4(w, w, w, w));
            cx.flushed = cx.pos & ~15u;
            cx.ring_lo = lo;
            __syncthreads();
            return;
        }
        cx_bulk_match(cx, off, ml, tid);
        cx.pos += ml;
    }
    cx_reload(cx, tid);
}

/*
 * Execute a chunk.  Thread t offers sequence t (valid: t < n, n <= CX_THREADS uniform):
 *   ll, ml, off   literal length, match length (0: none -- the last sequence of an LZ4 block, or
 *                 the literal tail of a zstd block), match distance
 *   lit           where the ll literal bytes are (global or shared memory); rle_byte >= 0: the
 *                 literals are ll copies of that byte
 *   cum           inclusive prefix sum of ll + ml over the chunk (cx_scan2)
 *   pre           a status the front end already found for this sequence (ST_OK: none)
 * Returns how many sequences were executed (a prefix of the chunk; the caller offers the rest
 * again), 0 when the block failed (cx.err).  A sequence with a run of CX_BIG bytes or more is
 * executed alone, as a bulk operation.
 */
CRYO_DEV uint32_t cx_chunk(Cx &cx, uint32_t n, uint32_t ll, uint32_t ml, uint32_t off, const uint8_t *lit,
                           int rle_byte, uint32_t cum, int pre, uint32_t tid)
{
    CxSh *sh = cx.sh;
    const uint32_t lane = tid & 31u, warp = tid >> 5;
    const bool     valid = tid < n;
    const uint32_t pos0 = cx.pos;
    const uint32_t end = pos0 + cum, start = end - ll - ml, mpos = start + ll;
    const bool     big = valid && (ll >= CX_BIG || ml >= CX_BIG);

    /* the prefix that is executed now: up to the first big sequence or the first one past the span */
    if (tid == 0)
    {
        sh->cut = n;
        sh->err = ~0u;
        sh->njobs = 0;
    }
    __syncthreads();
    {
        const uint32_t stop = __ballot_sync(CRYO_FULL, !valid || big || cum > CX_SPAN);

        if (stop && lane == 0)
            atomicMin(&sh->cut, (warp << 5) + (uint32_t) __ffs((int) stop) - 1u);
    }
    __syncthreads();
    uint32_t       k = sh->cut;
    const bool     alone = k == 0;      /* sequence 0 is big (n >= 1 and cum_0 < 2 * CX_BIG <= CX_SPAN otherwise) */

    if (alone)
        k = 1;
    /* checks of the sequences that are executed now: the first failing one names the status */
    {
        int e = ST_OK;

        if (tid < k)
        {
            if (pre != ST_OK)
                e = pre;
            else if (end > cx.cap || end < pos0)
                e = ST_OUTPUT;
            else if (ml && (off == 0 || off > mpos))
                e = ST_OFFSET;
        }
        const uint32_t bad = __ballot_sync(CRYO_FULL, e != ST_OK);

        if (bad && lane == (uint32_t) __ffs((int) bad) - 1u)
            atomicMin(&sh->err, (tid << 8) | (uint32_t) e);
    }
    /* runs for a whole warp are posted as jobs (not in a chunk that is one bulk operation) */
    bool own_lit = false, own_match = false;    /* ... unless the list is full: then the owning warp moves them */

    if (!alone && tid < k)
    {
        if (ll > CX_LANE)
        {
            const uint32_t slot = atomicAdd(&sh->njobs, 1u);

            if (slot < CX_JOBS)
            {
                CxJob &j = sh->jobs[slot];

                j.pos = start;
                j.len = ll;
                j.x = rle_byte >= 0 ? (uint32_t) rle_byte : ~0u;
                j.kind = 0;
                j.p = lit;
            }
            else
                own_lit = true;
        }
        if (ml > CX_LANE && ml < CX_BIG)
        {
            const uint32_t slot = atomicAdd(&sh->njobs, 1u);

            if (slot < CX_JOBS)
            {
                CxJob &j = sh->jobs[slot];

                j.pos = mpos;
                j.len = ml;
                j.x = off;
                j.kind = 1;
                j.p = nullptr;
            }
            else
                own_match = true;
        }
    }
    __syncthreads();
    if (sh->err != ~0u)
    {
        cx.err = (int) (sh->err & 0xFFu);
        return 0;
    }
    if (alone)
    {
        /* sequence 0 of thread 0 becomes uniform through shared memory */
        unsigned long long *slot = reinterpret_cast<unsigned long long *>(sh->wsa);

        if (tid == 0)
        {
            sh->bcast[0] = ll;
            sh->bcast[1] = ml;
            sh->bcast[2] = off;
            sh->bcast[3] = (uint32_t) rle_byte;
            slot[0] = (unsigned long long) (uintptr_t) lit;
        }
        __syncthreads();
        const uint32_t ll0 = sh->bcast[0], ml0 = sh->bcast[1], off0 = sh->bcast[2];
        const int      rb0 = (int) sh->bcast[3];
        const uint8_t *lit0 = reinterpret_cast<const uint8_t *>((uintptr_t) slot[0]);

        CXP(8)
        cx_bulk(cx, ll0, lit0, rb0, ml0, off0, tid);
        CXP(13)
        CXP_COUNT(18, 1)
        return 1;
    }
    CXP(8)
    const bool mine = tid < k;
    /* positions below lo are not in the ring any more (or never were, after a bulk operation) */
    const uint32_t kend = pos0 + CX_SPAN;
    const uint32_t lo = kend > CX_RING && kend - CX_RING > cx.ring_lo ? kend - CX_RING : cx.ring_lo;

    /* ---- literals: no dependencies; every byte of the chunk starts out final ---- */
    if (tid == k - 1u)
        sh->bcast[0] = end;
    for (uint32_t a = 8u * tid; a < CX_SPAN; a += 8u * CX_THREADS)
        st16(reinterpret_cast<uint8_t *>(cx.dl + a), make_uint4(0u, 0u, 0u, 0u));
    if (mine && ll && ll <= CX_LANE)
    {
        if (rle_byte >= 0)
            for (uint32_t i = 0; i < ll; i++)
                cx.ring[(start + i) & CX_RMASK] = (uint8_t) rle_byte;
        else
            for (uint32_t i = 0; i < ll; i++)
                cx.ring[(start + i) & CX_RMASK] = lit[i];
    }
    const uint32_t njobs = sh->njobs < CX_JOBS ? sh->njobs : CX_JOBS;

    for (uint32_t q = warp; q < njobs; q += CX_WARPS)
    {
        const CxJob &j = sh->jobs[q];

        if (j.kind != 0)
            continue;
        if (j.x != ~0u)
            cx_ring_fill(cx, j.pos, (uint8_t) j.x, j.len, lane, 32);
        else
            cx_ring_put(cx, j.pos, j.p, j.len, lane, 32);
    }
    for (uint32_t m = __ballot_sync(CRYO_FULL, own_lit); m; m &= m - 1)
    {
        const int      j = __ffs((int) m) - 1;
        const uint32_t jl = __shfl_sync(CRYO_FULL, ll, j), js = __shfl_sync(CRYO_FULL, start, j);
        const int      jr = __shfl_sync(CRYO_FULL, rle_byte, j);
        const uint8_t *jp = reinterpret_cast<const uint8_t *>(
            (uintptr_t) __shfl_sync(CRYO_FULL, (unsigned long long) (uintptr_t) lit, j));

        if (jr >= 0)
            cx_ring_fill(cx, js, (uint8_t) jr, jl, lane, 32);
        else
            cx_ring_put(cx, js, jp, jl, lane, 32);
    }
    __syncthreads();                    /* dl[] cleared, bcast */
    CXP(9)
    const uint32_t kpos = sh->bcast[0];         /* end of the chunk's last sequence */
    /*
     * ---- matches, step 1: every match byte records the distance to its source (16 bits: a chunk covers
     * CX_SPAN bytes and the window of an LZ4 block is 64 KiB; a zstd match from further back than CX_FAR
     * is copied here instead -- its source is older than the chunk, so it is final).  Stores only, no
     * loads.  A match that overlaps itself (offset < length) repeats the offset bytes before it: byte i
     * is sourced from byte i mod offset of that period, so the chain inside the match is cut here, not
     * by the rounds below. ----
     */
    if (mine && ml && ml <= CX_LANE)
    {
        if (off > CX_FAR)
        {
            const uint32_t s0 = mpos - off;

            for (uint32_t i = 0; i < ml; i++)
                cx.ring[(mpos + i) & CX_RMASK] = cx_src(cx, s0 + i, lo);
        }
        else
        {
            uint16_t *dp = cx.dl + (mpos - pos0);
            uint32_t  r = 0;

            for (uint32_t i = 0; i < ml; i++)
            {
                dp[i] = (uint16_t) (off + i - r);
                r++;
                r = r == off ? 0u : r;
            }
        }
    }
    /* the longer ones: a warp each, from the job list, then what did not fit it */
    for (uint32_t q = warp;; q += CX_WARPS)
    {
        uint32_t jm, jo, jp;

        if (q < njobs)
        {
            const CxJob &j = sh->jobs[q];

            if (j.kind != 1)
                continue;
            jm = j.len;
            jo = j.x;
            jp = j.pos;
        }
        else
        {
            /* the owning warp's own: one per pass of this loop */
            const uint32_t m = __ballot_sync(CRYO_FULL, own_match);

            if (m == 0)
                break;
            const int j = __ffs((int) m) - 1;

            jm = __shfl_sync(CRYO_FULL, ml, j);
            jo = __shfl_sync(CRYO_FULL, off, j);
            jp = __shfl_sync(CRYO_FULL, mpos, j);
            if ((int) lane == j)
                own_match = false;
        }
        if (jo > CX_FAR)
        {
            for (uint32_t i = lane; i < jm; i += 32)
                cx.ring[(jp + i) & CX_RMASK] = cx_src(cx, jp - jo + i, lo);
            continue;
        }
        const uint32_t step = jo >= 32u ? 32u : 32u % jo;
        uint32_t       r = lane < jo ? lane : lane % jo;
        uint16_t      *dp = cx.dl + (jp - pos0);

        for (uint32_t i = lane; i < jm; i += 32)
        {
            dp[i] = (uint16_t) (jo + i - r);
            r += step;
            r = r >= jo ? r - jo : r;
        }
    }
    __syncthreads();
    CXP(10)
    /*
     * ---- matches, step 2: pointer jumping.  A warp owns tiles of 512 bytes of the chunk and takes them
     * 32 consecutive bytes at a time, one per lane: the bytes of one match then read consecutive
     * sources, so the scattered 1- and 2-byte accesses of a row fall into few shared-memory wavefronts
     * (with 16 consecutive bytes per thread every access of a warp went to 8 banks).  An unresolved
     * byte whose source lies before the chunk, or is final, takes its value; one whose source is
     * unresolved adds that byte's distance to its own.  A round needs no barrier between its reads
     * and its writes: a byte that reads a distance another thread is replacing sees the old or the
     * new one, and both name a byte with the value it is after; a value is written (and fenced)
     * before its distance is cleared, and read after the distance was seen cleared.  The barrier at
     * the end of a round only tells whether any byte is left (bar.red.or).
     */
    {
        /* rows of 32 bytes are dealt out to the warps in turn (row R to warp R mod CX_WARPS), so that a chunk of any
         * size keeps every warp busy: with one 512-byte tile per warp a 9 KB chunk of a dense block left 14 of 32 warps
         * idle.  Bit b of rows[which]: row warp + CX_WARPS * (16 * which + b) has unresolved bytes. */
        const uint32_t nrows = (kpos - pos0 + 31u) >> 5;
        uint32_t       rows[CX_TPW];

#pragma unroll
        for (uint32_t which = 0; which < CX_TPW; which++)
        {
            uint32_t m = 0;

#pragma unroll
            for (uint32_t b = 0; b < 16; b++)
                if (warp + CX_WARPS * (16u * which + b) < nrows)
                    m |= 1u << b;
            rows[which] = m;
        }

        for (;;)
        {
            CXP_COUNT(16, 1)
            uint32_t any = 0;

            /* two jumps per barrier: the second one sees whatever the other warps have written meanwhile (old or
             * new, both valid), and a quarter of the executor's time went into waiting at the round barrier */
#pragma unroll 1
            for (uint32_t pass = 0; pass < 2u; pass++)
            {
#pragma unroll
            for (uint32_t which = 0; which < CX_TPW; which++)
            {
                uint32_t       r = rows[which];
                const uint32_t base = ((warp + CX_WARPS * 16u * which) << 5) + lane;    /* byte of this lane in row slot 0 */

                if (r == 0)
                    continue;
#pragma unroll
                for (uint32_t h = 0; h < 16; h += 4)
                {
                    if (((r >> h) & 15u) == 0u)
                        continue;
                    uint32_t d[4], ds[4];
                    uint8_t  v[4];

#pragma unroll
                    for (uint32_t q = 0; q < 4; q++)
                        d[q] = cx.dl[base + CX_ROWSTEP * (h + q)];
#pragma unroll
                    for (uint32_t q = 0; q < 4; q++)
                    {
                        const uint32_t idx = base + CX_ROWSTEP * (h + q);

                        ds[q] = (d[q] != 0u && d[q] <= idx) ? (uint32_t) cx.dl[idx - d[q]] : 0u;
                    }
                    CX_COMPILER_FENCE();        /* the values are read after the distances */
#pragma unroll
                    for (uint32_t q = 0; q < 4; q++)
                        v[q] = (d[q] != 0u && ds[q] == 0u) ? cx_src(cx, pos0 + base + CX_ROWSTEP * (h + q) - d[q], lo) : (uint8_t) 0;
#pragma unroll
                    for (uint32_t q = 0; q < 4; q++)
                        if (d[q] != 0u && ds[q] == 0u)
                            cx.ring[(pos0 + base + CX_ROWSTEP * (h + q)) & CX_RMASK] = v[q];
                    __threadfence_block();      /* the values before the cleared distances */
#pragma unroll
                    for (uint32_t q = 0; q < 4; q++)
                    {
                        if (d[q] != 0u)
                            cx.dl[base + CX_ROWSTEP * (h + q)] = (uint16_t) (ds[q] == 0u ? 0u : d[q] + ds[q]);
                        if (__ballot_sync(CRYO_FULL, d[q] != 0u && ds[q] != 0u) == 0u)
                            r &= ~(1u << (h + q));
                    }
                }
                rows[which] = r;
                any = pass ? any | r : any;
            }
            }
            if (!CX_SYNC_OR(any != 0u))
                break;
        }
    }
    CXP(11)
    cx.pos = kpos;
    cx_drain(cx, tid);
    CXP(12)
    CXP_COUNT(17, 1)
    return k;
}

/* end of block: everything to global memory */
CRYO_DEV void cx_finish(Cx &cx, uint32_t tid)
{
    __syncthreads();
    cx_drain(cx, tid);
    cx_drain_tail(cx, tid);
    cx.flushed = cx.pos;
}
