/*
 * lz4_decode.cuh -- batched LZ4 *block* decompression, one CTA per cryo block.
 *
 * Replaces LZ4_decompress_safe as called by the reference at compression.c:84
 * (lz4_decompress, compression.c:79-91): raw LZ4 block in, at most `cap` bytes
 * out, negative result (here: a non-zero status) on any malformed input.  The
 * acceptance rules are LZ4_decompress_safe's: input consumed
 * exactly, last literal run / last match distance rules, offsets inside the
 * output; offset 0 is rejected.
 *
 * Structure: warp 0 parses the token stream out of a shared-memory window of
 * the compressed input (length-extension runs are scanned 32 bytes at a time
 * with a ballot), and drives cryo_exec.cuh, which owns the output tile and
 * dispatches long literal runs / long matches to the whole CTA.
 */
#pragma once
#include "cryo_exec.cuh"

#define LZ4D_THREADS 128
#define LZ4D_WIN     (8 * 1024)       /* bytes of compressed input staged per refill */

/* dynamic shared memory layout */
#define LZ4D_SM_EXEC   0
#define LZ4D_SM_TILE   128
#define LZ4D_SM_PAT    (LZ4D_SM_TILE + EX_TILE)
#define LZ4D_SM_WIN    (LZ4D_SM_PAT + EX_PAT_BYTES)
#define LZ4D_SMEM      (LZ4D_SM_WIN + LZ4D_WIN)

struct Lz4In
{
    const uint8_t *base;        /* 16-byte aligned address at or before the stream */
    uint8_t    *win;            /* shared window, mirrors base[wbase, wbase + LZ4D_WIN) */
    uint32_t    wbase;          /* multiple of 16 */
    uint32_t    end;            /* stream end, in `base` coordinates */
};

/* make base[ip, ip+need) readable through the window (need <= LZ4D_WIN - 16) */
CRYO_DEV void lz4_window(Lz4In &in, uint32_t ip, uint32_t need, uint32_t lane, bool force = false)
{
    if (!force && ip >= in.wbase && ip + need <= in.wbase + LZ4D_WIN)
        return;
    __syncwarp();
    in.wbase = align_down16(ip);
    uint32_t lim = (in.end + 15u) & ~15u;     /* never read past the last granule of the stream */

#pragma unroll 4
    for (uint32_t v = lane; v < LZ4D_WIN / 16; v += 32)
    {
        uint32_t a = in.wbase + 16 * v;

        if (a < lim)
            st16(in.win + 16 * v, ld16(in.base + a));
    }
    __syncwarp();
}

/* length extension: sum of bytes up to and including the first one != 255 */
CRYO_DEV uint32_t lz4_read_ext(Lz4In &in, uint32_t &ip, uint32_t lane, int &err)
{
    uint32_t add = 0;

    for (;;)
    {
        lz4_window(in, ip, 32, lane);
        uint32_t idx = ip + lane;
        uint32_t b = idx < in.end ? in.win[idx - in.wbase] : 0u;
        uint32_t m = __ballot_sync(CRYO_FULL, b != 255u);

        if (m == 0)
        {
            add += 255u * 32u;
            ip += 32;
            if (add > 0x40000000u)
            {
                err = ST_INPUT;
                return add;
            }
            continue;
        }
        uint32_t k = (uint32_t) __ffs((int) m) - 1u;

        add += 255u * k + __shfl_sync(CRYO_FULL, b, (int) k);
        ip += k + 1;
        if (ip > in.end)
            err = ST_INPUT;
        return add;
    }
}

/*
 * Decode one LZ4 block.  Called by every thread of the CTA; src may have any
 * alignment, out must be 16-byte aligned.
 */
CRYO_DEV void lz4_decode_block(const uint8_t *src, uint32_t csize, uint8_t *out, uint32_t cap,
                               uint32_t *out_size, int32_t *status)
{
    uint8_t    *smem = CRYO_SMEM_BASE();
    ExecShared *sh = reinterpret_cast<ExecShared *>(smem + LZ4D_SM_EXEC);
    uint8_t    *tile = smem + LZ4D_SM_TILE;
    uint8_t    *pat = smem + LZ4D_SM_PAT;
    const uint32_t tid = threadIdx.x, nthr = LZ4D_THREADS;

    if (tid >= 32)
    {
        exec_worker_loop(out, tile, pat, sh, tid, nthr);
        return;
    }

    /* ---- master warp ---- */
    Exec  e;
    Lz4In in;
    int   err = ST_OK;
    const uint32_t lane = tid;
    uint32_t delta = (uint32_t) ((uintptr_t) src & 15u);
    uint32_t ip = delta;

    exec_init(e, out, cap, tile, pat, sh);
    in.base = src - delta;
    in.win = smem + LZ4D_SM_WIN;
    in.end = csize + delta;
    in.wbase = 0;
    if (csize == 0)
        err = ST_INPUT;
    else
        lz4_window(in, ip, 1, lane, true);

    while (err == ST_OK)
    {
        lz4_window(in, ip, 1, lane);
        if (ip >= in.end)
        {
            err = ST_INPUT;
            break;
        }
        uint32_t token = in.win[ip - in.wbase];
        uint32_t ll = token >> 4;

        ip++;
        if (ll == 15)
        {
            ll += lz4_read_ext(in, ip, lane, err);
            if (err)
                break;
        }
        if (ip + ll > in.end || ip + ll < ip)
        {
            err = ST_INPUT;
            break;
        }
        if (e.pos + ll > cap || e.pos + ll < e.pos)
        {
            err = ST_OUTPUT;
            break;
        }
        bool last = (ip + ll == in.end);

        /* LZ4_decompress_safe: a literal run ending within 12 bytes of the output
         * capacity or within 8 bytes of the input end must be the last one */
        if (!last && (e.pos + ll + 12 > cap || ip + ll + 8 > in.end))
        {
            err = (ip + ll + 8 > in.end && e.pos + ll + 12 <= cap) ? ST_INPUT : ST_OUTPUT;
            break;
        }
        if (ll)
        {
            if (ll >= EX_BULK)
                exec_literals(e, in.base + ip, nullptr, ll, tid, nthr);
            else
            {
                lz4_window(in, ip, ll, lane);
                exec_literals(e, in.base + ip, in.win + (ip - in.wbase), ll, tid, nthr);
            }
            ip += ll;
        }
        if (last)
            break;
        lz4_window(in, ip, 2, lane);
        uint32_t off = in.win[ip - in.wbase] | ((uint32_t) in.win[ip + 1 - in.wbase] << 8);
        uint32_t ml = token & 15u;

        ip += 2;
        if (ml == 15)
        {
            ml += lz4_read_ext(in, ip, lane, err);
            if (err)
                break;
        }
        ml += 4;
        if (off == 0 || off > e.pos)
        {
            err = ST_OFFSET;
            break;
        }
        if (e.pos + ml + 5 > cap || e.pos + ml < e.pos)
        {
            err = ST_OUTPUT;
            break;
        }
        exec_match(e, off, ml, tid, nthr);
    }
    exec_finish(e, tid, nthr);
    if (tid == 0)
    {
        *out_size = err == ST_OK ? e.pos : 0u;
        *status = err;
    }
}
