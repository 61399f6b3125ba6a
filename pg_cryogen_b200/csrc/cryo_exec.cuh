/*
 * cryo_exec.cuh -- the output side shared by the LZ4 and zstd decoders: executing
 * (literal run, match) sequences into one cryo block of output.
 *
 * One CTA owns one output block.  Warp 0 (the "master") walks the sequence stream,
 * which is inherently serial, and appends short literal runs and short matches to
 * a shared-memory TILE that mirrors out[tbase, tbase+EX_TILE).  Matches whose
 * source lies in the tile are served from shared memory (tens of cycles) instead
 * of an L2 round trip; the tile keeps EX_HIST bytes of history when it slides.
 * Everything long -- the >1 MB zero run of a sparse block, long literal runs of a
 * dense one, full tiles -- is executed by the whole CTA with coalesced 16-byte
 * global stores through a small command protocol:
 *
 *     master: write cmd to shared memory; __syncthreads(); handle; __syncthreads()
 *     worker: loop { __syncthreads(); read cmd; EXIT? ; handle; __syncthreads() }
 *
 * Overlapping matches (offset < length, e.g. the zero run) are periodic fills:
 * the period is staged once in shared memory and every thread emits aligned
 * 16-byte vectors of it.
 */
#pragma once
#include "cryo_common.cuh"

#ifndef EX_TILE
#define EX_TILE   (24 * 1024)     /* bytes of output mirrored in shared memory */
#endif
#ifndef EX_HIST
#define EX_HIST   (8 * 1024)      /* history kept when the tile slides */
#endif
#define EX_BULK   1024u           /* runs at least this long go to the whole CTA */
#define EX_PAT_MAXOFF 2048u       /* periods below this are staged in shared memory */
#define EX_PAT_BYTES  (2 * EX_PAT_MAXOFF + 32)

enum { EXC_EXIT = 0, EXC_FLUSH = 1, EXC_SLIDE = 2, EXC_COPY = 3, EXC_MATCH = 4, EXC_FILLBYTE = 5 };

struct ExecShared
{
    /* command mailbox (written by lane 0 of the master warp) */
    int32_t     op;
    uint32_t    n;              /* bytes */
    uint32_t    off;            /* EXC_MATCH: offset; EXC_FILLBYTE: byte value */
    const uint8_t *src;         /* EXC_COPY: source (global) */
    /* tile state published with every command */
    uint32_t    tbase;
    uint32_t    flushed;
    uint32_t    pos;
    int32_t     pad;
};

/* master-side (warp-uniform, register-resident) view of the output */
struct Exec
{
    uint8_t    *out;            /* global output block, 16-byte aligned */
    uint8_t    *tile;           /* shared, EX_TILE bytes, 16-byte aligned */
    uint8_t    *pat;            /* shared, EX_PAT_BYTES, 16-byte aligned */
    ExecShared *sh;
    uint32_t    cap;            /* output capacity (block_size) */
    uint32_t    pos;            /* next output byte */
    uint32_t    tbase;          /* tile[i] mirrors out[tbase + i]; multiple of 16 */
    uint32_t    flushed;        /* out[0, flushed) is in global memory */
};

/* ---- CTA-wide pieces (every thread of the CTA calls these together) ---- */

/* write tile bytes [flushed, pos) to global memory */
CRYO_DEV void exec_flush_range(uint8_t *out, const uint8_t *tile, uint32_t tbase,
                               uint32_t flushed, uint32_t pos, uint32_t tid, uint32_t nthr)
{
    if (pos <= flushed)
        return;
    uint32_t f0 = align_down16(flushed);      /* >= tbase: tile holds the whole granule */
    uint32_t p0 = align_down16(pos);

    for (uint32_t a = f0 + 16 * tid; a < p0; a += 16 * nthr)
        st16(out + a, ld16(tile + (a - tbase)));
    if (p0 >= f0 && tid < pos - p0)
    {
        if (p0 + tid >= flushed)
            out[p0 + tid] = tile[p0 - tbase + tid];
    }
}

/*
 * Overlapping or plain match executed on global memory: out[pos+i] = out[pos+i-off].
 * All of out[0,pos) must already be in global memory and visible.
 * Contains __syncthreads(): call from all threads.
 */
CRYO_DEV void exec_bulk_match(uint8_t *out, uint8_t *pat, uint32_t pos, uint32_t off, uint32_t n,
                              uint32_t tid, uint32_t nthr)
{
    if (off >= n)
    {
        team_copy(out + pos, out + pos - off, n, tid, nthr);
        return;
    }
    if (off < EX_PAT_MAXOFF)
    {
        /* stage k whole periods (k*off >= EX_PAT_MAXOFF) plus wrap-around */
        uint32_t k = (EX_PAT_MAXOFF + off - 1) / off;
        uint32_t plen = k * off;
        const uint8_t *src = out + pos - off;

        if (off == 1)
        {
            team_fill_byte(out + pos, src[0], n, tid, nthr);
            return;
        }
        for (uint32_t j = tid; j < plen + 32; j += nthr)
            pat[j] = src[j % off];
        __syncthreads();
        team_fill_from_pattern(out + pos, pat, plen, 0, n, tid, nthr);
        return;
    }
    /* long period: out[pos-off, pos+done) is periodic; copy the largest whole
     * number of periods available each round (doubles every round) */
    uint32_t done = 0;

    while (done < n)
    {
        uint32_t avail = ((off + done) / off) * off;
        uint32_t m = n - done < avail ? n - done : avail;

        team_copy(out + pos + done, out + pos + done - avail, m, tid, nthr);
        done += m;
        __syncthreads();
    }
}

/* one command, executed by all threads between the two barriers of the protocol */
CRYO_DEV void exec_handle(uint8_t *out, uint8_t *tile, uint8_t *pat, const ExecShared *sh,
                          uint32_t tid, uint32_t nthr)
{
    int32_t  op = sh->op;
    uint32_t n = sh->n, off = sh->off;
    uint32_t tbase = sh->tbase, flushed = sh->flushed, pos = sh->pos;
    const uint8_t *src = sh->src;

    exec_flush_range(out, tile, tbase, flushed, pos, tid, nthr);
    switch (op)
    {
        case EXC_SLIDE:
        {
            /* keep the last EX_HIST bytes: tile[EX_TILE-EX_HIST, EX_TILE) -> tile[0, EX_HIST) */
            uint4 r[8];             /* nthr >= 64 so 8 vectors per thread cover EX_HIST */

#pragma unroll
            for (uint32_t i = 0; i < 8; i++)
            {
                uint32_t v = tid + i * nthr;

                if (v < EX_HIST / 16)
                    r[i] = ld16(tile + (EX_TILE - EX_HIST) + 16 * v);
            }
            __syncthreads();
#pragma unroll
            for (uint32_t i = 0; i < 8; i++)
            {
                uint32_t v = tid + i * nthr;

                if (v < EX_HIST / 16)
                    st16(tile + 16 * v, r[i]);
            }
            break;
        }
        case EXC_COPY:
            team_copy(out + pos, src, n, tid, nthr);
            break;
        case EXC_MATCH:
            __syncthreads();            /* the flush above must be visible to the reads below */
            exec_bulk_match(out, pat, pos, off, n, tid, nthr);
            break;
        case EXC_FILLBYTE:
            team_fill_byte(out + pos, (uint8_t) off, n, tid, nthr);
            break;
        default:
            break;
    }
}

/* worker warps: serve commands until EXC_EXIT */
CRYO_DEV void exec_worker_loop(uint8_t *out, uint8_t *tile, uint8_t *pat, const ExecShared *sh,
                               uint32_t tid, uint32_t nthr)
{
    for (;;)
    {
        __syncthreads();
        if (sh->op == EXC_EXIT)
            break;
        exec_handle(out, tile, pat, sh, tid, nthr);
        __syncthreads();
    }
}

/* ---- master-side (warp 0, all 32 lanes, warp-uniform arguments) ---- */

CRYO_DEV void exec_init(Exec &e, uint8_t *out, uint32_t cap, uint8_t *tile, uint8_t *pat,
                        ExecShared *sh)
{
    e.out = out;
    e.cap = cap;
    e.tile = tile;
    e.pat = pat;
    e.sh = sh;
    e.pos = 0;
    e.tbase = 0;
    e.flushed = 0;
}

CRYO_DEV void exec_issue(Exec &e, int32_t op, uint32_t n, uint32_t off, const uint8_t *src,
                         uint32_t tid, uint32_t nthr)
{
    if (tid == 0)
    {
        e.sh->op = op;
        e.sh->n = n;
        e.sh->off = off;
        e.sh->src = src;
        e.sh->tbase = e.tbase;
        e.sh->flushed = e.flushed;
        e.sh->pos = e.pos;
    }
    __syncthreads();
    if (op != EXC_EXIT)
    {
        exec_handle(e.out, e.tile, e.pat, e.sh, tid, nthr);
        __syncthreads();
    }
}

/* after a bulk operation wrote n bytes at pos directly to global memory */
CRYO_DEV void exec_after_bulk(Exec &e, uint32_t n, uint32_t lane)
{
    e.pos += n;
    e.flushed = e.pos;
    e.tbase = align_down16(e.pos);
    if (lane < e.pos - e.tbase)
        e.tile[lane] = e.out[e.tbase + lane];
    __syncwarp();
}

/* make room: called when the tile is full (pos == tbase + EX_TILE) */
CRYO_DEV void exec_slide(Exec &e, uint32_t tid, uint32_t nthr)
{
    exec_issue(e, EXC_SLIDE, 0, 0, nullptr, tid, nthr);
    e.flushed = e.pos;
    e.tbase = e.pos - EX_HIST;
}

/* literal run from a byte source readable by warp 0 (shared or global memory) */
CRYO_DEV void exec_literals_small(Exec &e, const uint8_t *src, uint32_t n, uint32_t tid,
                                  uint32_t nthr)
{
    while (n)
    {
        uint32_t room = e.tbase + EX_TILE - e.pos;

        if (room == 0)
        {
            exec_slide(e, tid, nthr);
            continue;
        }
        uint32_t m = n < room ? n : room;
        uint8_t *d = e.tile + (e.pos - e.tbase);

        for (uint32_t i = tid; i < m; i += 32)
            d[i] = src[i];
        __syncwarp();
        e.pos += m;
        src += m;
        n -= m;
    }
}

/* literal run from global memory, any length */
CRYO_DEV void exec_literals(Exec &e, const uint8_t *gsrc, const uint8_t *ssrc, uint32_t n,
                            uint32_t tid, uint32_t nthr)
{
    if (n >= EX_BULK)
    {
        exec_issue(e, EXC_COPY, n, 0, gsrc, tid, nthr);
        exec_after_bulk(e, n, tid);
    }
    else
        exec_literals_small(e, ssrc ? ssrc : gsrc, n, tid, nthr);
}

/* short run of one byte value, through the tile */
CRYO_DEV void exec_fill_small(Exec &e, uint8_t b, uint32_t n, uint32_t tid, uint32_t nthr)
{
    while (n)
    {
        uint32_t room = e.tbase + EX_TILE - e.pos;

        if (room == 0)
        {
            exec_slide(e, tid, nthr);
            continue;
        }
        uint32_t m = n < room ? n : room;
        uint8_t *d = e.tile + (e.pos - e.tbase);

        for (uint32_t i = tid; i < m; i += 32)
            d[i] = b;
        __syncwarp();
        e.pos += m;
        n -= m;
    }
}

CRYO_DEV void exec_fill_byte(Exec &e, uint8_t b, uint32_t n, uint32_t tid, uint32_t nthr)
{
    exec_issue(e, EXC_FILLBYTE, n, b, nullptr, tid, nthr);
    exec_after_bulk(e, n, tid);
}

/* match: out[pos+i] = out[pos+i-off], i < n.  Caller has validated off and bounds. */
CRYO_DEV void exec_match(Exec &e, uint32_t off, uint32_t n, uint32_t tid, uint32_t nthr)
{
    if (n >= EX_BULK)
    {
        exec_issue(e, EXC_MATCH, n, off, nullptr, tid, nthr);
        exec_after_bulk(e, n, tid);
        return;
    }
    while (n)
    {
        uint32_t room = e.tbase + EX_TILE - e.pos;

        if (room == 0)
        {
            exec_slide(e, tid, nthr);
            continue;
        }
        uint32_t m = n < room ? n : room;
        uint32_t src = e.pos - off;
        uint8_t *d = e.tile + (e.pos - e.tbase);

        if (off >= 32 || off >= m)
        {
            /* a 32-byte step never reads what the same step writes */
            for (uint32_t i0 = 0; i0 < m; i0 += 32)
            {
                uint32_t i = i0 + tid;

                if (i < m)
                {
                    uint32_t s = src + i;

                    d[i] = s >= e.tbase ? e.tile[s - e.tbase] : e.out[s];
                }
                if (off < m)
                    __syncwarp();
            }
        }
        else
        {
            /* short period: every byte comes from the off bytes before pos */
            for (uint32_t i = tid; i < m; i += 32)
            {
                uint32_t s = src + (i % off);

                d[i] = s >= e.tbase ? e.tile[s - e.tbase] : e.out[s];
            }
        }
        __syncwarp();
        e.pos += m;
        n -= m;
    }
}

/* end of block: flush what is left and release the workers */
CRYO_DEV void exec_finish(Exec &e, uint32_t tid, uint32_t nthr)
{
    exec_issue(e, EXC_FLUSH, 0, 0, nullptr, tid, nthr);
    e.flushed = e.pos;
    exec_issue(e, EXC_EXIT, 0, 0, nullptr, tid, nthr);
}
