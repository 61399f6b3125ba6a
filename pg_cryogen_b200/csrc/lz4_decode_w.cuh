/*
 * lz4_decode_w.cuh -- batched LZ4 block decompression, ONE WARP per cryo block
 * (the throughput path; lz4_decode_c.cuh is the one-CTA-per-block decoder).
 *
 * Replaces LZ4_decompress_safe as called at reference compression.c:84, with
 * the same acceptance rules.  The token stream is parsed out of
 * a per-warp shared-memory window of the compressed input, refilled with
 * coalesced 16-byte loads; output goes through cryo_wexec.cuh.  A sequence
 * whose literal and match lengths need no extension bytes -- almost all of
 * them in heap-tuple data -- takes a straight-line fast path: two dependent
 * shared-memory reads (token, offset), one predicated literal move, one
 * predicated match move.
 *
 * A warp takes about half a microsecond per sequence, so a block with a hundred thousand of
 * them is not for this decoder.  How many sequences a block has is not in the format, and
 * sampling the stream does not tell (a literal-heavy block looks dense wherever a sample starts
 * but at a true token; the item-id array at the head of a cryo block is dense whatever follows):
 * the warp simply counts, and past `budget` sequences it gives the block up -- writes nothing
 * to status -- and the caller queues it for the CTA decoder, which starts it from scratch.
 */
#pragma once
#include "cryo_wexec.cuh"

#define LZ4W_WARPS   2
#define LZ4W_THREADS (32 * LZ4W_WARPS)
#define LZ4W_WIN     4096u
#define LZ4W_PER_WARP (WX_RING + LZ4W_WIN)
#define LZ4W_SMEM    (LZ4W_WARPS * LZ4W_PER_WARP)

struct Lz4Win
{
    const uint8_t *base;        /* 16-byte aligned address at or before the stream */
    uint8_t    *win;            /* shared window: base[wbase, wbase + LZ4W_WIN) */
    uint32_t    wbase;
    uint32_t    end;            /* stream end in `base` coordinates */
};

CRYO_DEV void lz4w_refill(Lz4Win &in, uint32_t ip, uint32_t lane)
{
    __syncwarp();
    in.wbase = ip & ~15u;
    const uint32_t lim = (in.end + 15u) & ~15u;

#pragma unroll
    for (uint32_t k = 0; k < LZ4W_WIN / 512; k++)
    {
        uint32_t a = in.wbase + 512u * k + 16u * lane;

        if (a < lim)
            st16(in.win + 512u * k + 16u * lane, ld16(in.base + a));
    }
    __syncwarp();
}

/* make base[ip, ip+need) readable through the window; need <= LZ4W_WIN - 16 */
CRYO_DEV void lz4w_need(Lz4Win &in, uint32_t ip, uint32_t need, uint32_t lane)
{
    if (ip < in.wbase || ip + need > in.wbase + LZ4W_WIN)
        lz4w_refill(in, ip, lane);
}

CRYO_DEV uint32_t lz4w_read_ext(Lz4Win &in, uint32_t &ip, uint32_t lane, int &err)
{
    uint32_t add = 0;

    for (;;)
    {
        lz4w_need(in, ip, 32, lane);
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

/* one warp decodes one block; `smem` is this warp's LZ4W_PER_WARP bytes.  budget: sequences after which the
 * block is given up (0: never).  Returns true when it was. */
CRYO_DEV bool lz4w_decode_block(const uint8_t *src, uint32_t csize, uint8_t *out, uint32_t cap,
                                uint32_t *out_size, int32_t *status, uint8_t *smem, uint32_t lane, uint32_t budget = 0)
{
    uint32_t nseq = 0;

    WOut   o;
    Lz4Win in;
    int    err = ST_OK;
    const uint32_t delta = (uint32_t) ((uintptr_t) src & 15u);
    uint32_t ip = delta;

    wx_init(o, out, cap, smem);
    in.base = src - delta;
    in.win = smem + WX_RING;
    in.end = csize + delta;
    in.wbase = 0;
    if (csize == 0)
        err = ST_INPUT;
    else
        lz4w_refill(in, ip, lane);

    /* bytes the window must hold ahead of ip at the top of an iteration: token, one
     * extension byte, a literal run below WX_BULK, the offset, one extension byte */
    const uint32_t AHEAD = WX_BULK + 8u;
    const uint32_t lim = (in.end + 15u) & ~15u;

    while (err == ST_OK)
    {
        if (ip + AHEAD > in.wbase + LZ4W_WIN && in.wbase + LZ4W_WIN < lim)
            lz4w_refill(in, ip, lane);
        if (ip >= in.end)
        {
            err = ST_INPUT;
            break;
        }
        if (budget && ++nseq > budget)
            return true;
        uint32_t rel = ip - in.wbase;
        uint32_t token = in.win[rel];
        uint32_t ll = token >> 4, ml = token & 15u;

        /*
         * Fast path: at most one length-extension byte each, runs up to 64 bytes, everything
         * well inside the input window and the output capacity, the match source inside the
         * ring and not overlapping its destination.  Two predicated moves each for literals
         * and match.
         */
        if (rel + 144u <= LZ4W_WIN)
        {
            uint32_t q = rel + 1, fl = ll, fm = ml;
            bool     ok = true;

            if (fl == 15u)
            {
                const uint32_t e = in.win[q++];

                fl += e;
                ok = e != 255u;
            }
            const uint32_t lit = q;

            q += fl;
            if (ok && fl <= 64u)
            {
                const uint32_t off = in.win[q] | ((uint32_t) in.win[q + 1] << 8);

                q += 2;
                if (fm == 15u)
                {
                    const uint32_t e = in.win[q++];

                    fm += e;
                    ok = e != 255u;
                }
                const uint32_t mlen = fm + 4u, mpos = o.pos + fl;
                const uint32_t used = q - rel;

                if (ok && mlen <= 64u && ip + used + 8u <= in.end && mpos + mlen + 21u <= cap &&
                    off >= mlen && off <= WX_RING - 64u && off <= mpos && mpos - off >= o.lo)
                {
                    if (lane < fl)
                        o.ring[(o.pos + lane) & WX_RMASK] = in.win[lit + lane];
                    if (lane + 32u < fl)
                        o.ring[(o.pos + lane + 32u) & WX_RMASK] = in.win[lit + lane + 32u];
                    __syncwarp();
                    if (lane < mlen)
                        o.ring[(mpos + lane) & WX_RMASK] = o.ring[(mpos - off + lane) & WX_RMASK];
                    if (lane + 32u < mlen)
                        o.ring[(mpos + lane + 32u) & WX_RMASK] = o.ring[(mpos - off + lane + 32u) & WX_RMASK];
                    o.pos = mpos + mlen;
                    ip += used;
                    __syncwarp();
                    if (o.pos - o.flushed >= WX_DRAIN)
                        wx_drain(o, lane);
                    continue;
                }
            }
        }
        ip++;
        if (ll == 15)
        {
            /* one extension byte covers runs up to 269; longer ones take the ballot scan */
            uint32_t b = ip < in.end ? in.win[rel + 1] : 255u;

            if (b != 255u)
            {
                ll += b;
                ip++;
            }
            else
            {
                ll += lz4w_read_ext(in, ip, lane, err);
                if (err)
                    break;
            }
        }
        if (ip + ll > in.end || ip + ll < ip)
        {
            err = ST_INPUT;
            break;
        }
        if (o.pos + ll > cap || o.pos + ll < o.pos)
        {
            err = ST_OUTPUT;
            break;
        }
        const bool last = (ip + ll == in.end);

        /* LZ4_decompress_safe: a literal run ending within 12 bytes of the output
         * capacity or within 8 bytes of the input end must be the last one */
        if (!last && (o.pos + ll + 12 > cap || ip + ll + 8 > in.end))
        {
            err = (ip + ll + 8 > in.end && o.pos + ll + 12 <= cap) ? ST_INPUT : ST_OUTPUT;
            break;
        }
        if (ll >= WX_BULK)
        {
            wx_literals(o, in.base + ip, ll, lane);
            ip += ll;
            if (last)
                break;
            lz4w_need(in, ip, 3, lane);
        }
        else
        {
            if (ip < in.wbase || ip + ll + 3 > in.wbase + LZ4W_WIN)
                lz4w_need(in, ip, ll + 3, lane);      /* only after a long extension scan */
            const uint8_t *lp = in.win + (ip - in.wbase);

            for (uint32_t i = lane; i < ll; i += 32)
                o.ring[(o.pos + i) & WX_RMASK] = lp[i];
            o.pos += ll;
            ip += ll;
            if (last)
                break;
        }
        rel = ip - in.wbase;
        uint32_t off = in.win[rel] | ((uint32_t) in.win[rel + 1] << 8);

        ip += 2;
        if (ml == 15)
        {
            uint32_t b = ip < in.end ? in.win[rel + 2] : 255u;

            if (b != 255u)
            {
                ml += b;
                ip++;
            }
            else
            {
                ml += lz4w_read_ext(in, ip, lane, err);
                if (err)
                    break;
            }
        }
        ml += 4;
        if (off == 0 || off > o.pos)
        {
            err = ST_OFFSET;
            break;
        }
        if (o.pos + ml + 5 > cap || o.pos + ml < o.pos)
        {
            err = ST_OUTPUT;
            break;
        }
        __syncwarp();               /* the literal bytes above may be this match's source */
        wx_match(o, off, ml, lane);
    }
    wx_drain_all(o, lane);
    if (lane == 0)
    {
        *out_size = err == ST_OK ? o.pos : 0u;
        *status = err;
    }
    return false;
}
