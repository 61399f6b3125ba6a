/*
 * cryo_pages.cuh -- the PostgreSQL page chain on either side of the codec.
 *
 * On disk a compressed cryo block is cut into 8 KiB PostgreSQL pages: the first carries a
 * CryoFirstPageHeader (48 bytes: compression_method, compressed_size, npages, created_xid), the
 * others a CryoPageHeader (32 bytes); every header names the chain's first page and the next one
 * (reference storage.h:26-67).  The reference's reader copies the payloads into one palloc'ed
 * buffer on the host before it calls cryo_decompress (cache.c:151-176), and its writer copies the
 * compressed bytes out page by page after cryo_compress (pg_cryogen.c:761-805).  Here both copies
 * happen in HBM: whole pages cross the bus as they lie in the buffer pool, and
 *
 *   gather  one CTA per cryo block reads the first page's header, checks the chain the host walked
 *           (first / next of every page, page count against compressed_size) and packs the payloads
 *           into the contiguous stream the decoders take -- 16-byte vectors throughout: the payload
 *           offsets 48 / 32 and sizes 8 144 / 8 160 are all multiples of 16;
 *   split   one CTA per cryo block writes the page images the reference would: zeroed page, header
 *           fields, payload; cryo_pages_needed pages.
 */
#pragma once
#include "cryo_common.cuh"

#define PG_PAGE         8192u
#define PG_HDR          32u             /* sizeof(CryoPageHeader) */
#define PG_HDR_FIRST    48u             /* sizeof(CryoFirstPageHeader) */
#define PG_OFF_LOWER    12u
#define PG_OFF_UPPER    14u
#define PG_OFF_SPECIAL  16u
#define PG_OFF_FIRST    24u
#define PG_OFF_NEXT     28u
#define PG_OFF_XID      32u
#define PG_OFF_METHOD   36u
#define PG_OFF_CSIZE    40u
#define PG_OFF_NPAGES   44u
#define PG_INVALID      0xFFFFFFFFu

/* per-block outcome of the gather, merged into the decode status afterwards */
#define PG_ST_OK            0
#define PG_ST_EMPTY         8           /* CRYOGPU_ST_EMPTY_BLOCK */
#define PG_ST_WRONG_START   9           /* CRYOGPU_ST_WRONG_START */
#define PG_ST_CHAIN         10          /* CRYOGPU_ST_CHAIN */
#define PG_METHOD_SKIP      0x7FFFFFFF  /* method handed to the decoders for a block whose chain failed */

#ifdef CRYO_EMU
#define PG_HOSTDEV static inline
#else
#define PG_HOSTDEV __host__ __device__ __forceinline__
#endif
PG_HOSTDEV uint32_t pg_pages_needed(uint64_t size)
{
    /* pg_cryogen.c:692-704 */
    return size <= PG_PAGE - PG_HDR_FIRST ? 1u
                                          : 1u + (uint32_t) ((size - (PG_PAGE - PG_HDR_FIRST) + (PG_PAGE - PG_HDR) - 1) / (PG_PAGE - PG_HDR));
}

/*
 * Gather, one CTA, one cryo block.  Chain entries [c0, c1): slot[e] = index of the page in `pages`,
 * blkno[e] = its block number in the relation.  The payloads go to comp + c0 * PG_PAGE (PG_PAGE bytes
 * per chain entry are reserved there).  Outputs, one per block: where and how long the stream is and the
 * method for the decoders (PG_METHOD_SKIP when the chain failed), the method as the header has it, and the
 * outcome of the gather.
 */
CRYO_DEV void pg_gather_block(const uint8_t *pages, const uint32_t *slot, const uint32_t *blkno, uint32_t c0, uint32_t c1,
                              uint8_t *comp, uint64_t *src_off, uint32_t *src_size, int32_t *dec_method,
                              int32_t *hdr_method, int32_t *chain_status, uint32_t max_csize, uint32_t tid, uint32_t nthr)
{
    int      st = PG_ST_OK;
    uint32_t csize = 0, method = 0;

    if (c1 <= c0)
        st = PG_ST_EMPTY;
    else
    {
        const uint8_t *p0 = pages + (size_t) slot[c0] * PG_PAGE;

        if ((ld4(p0 + 12) >> 16) == 0)          /* pd_lower | pd_upper << 16 */
            st = PG_ST_EMPTY;                   /* PageIsNew, cache.c:115-119 */
        else if (ld4(p0 + PG_OFF_FIRST) != blkno[c0])
            st = PG_ST_WRONG_START;             /* cache.c:125-129 */
        else
        {
            method = ld4(p0 + PG_OFF_METHOD);
            csize = ld4(p0 + PG_OFF_CSIZE);
            /* the chain the host walked must be long enough for compressed_size (the reference stops at an
             * invalid next and then fails in cryo_decompress, cache.c:163-178) */
            if (csize > max_csize || pg_pages_needed(csize) > c1 - c0)
                st = PG_ST_CHAIN;
        }
    }
    if (st == PG_ST_OK)
    {
        const uint32_t np = pg_pages_needed(csize);
        uint8_t *dst = comp + (size_t) c0 * PG_PAGE;
        uint32_t left = csize;

        for (uint32_t k = 0; k < np; k++)
        {
            const uint8_t *pg = pages + (size_t) slot[c0 + k] * PG_PAGE;
            const uint32_t hdr = k == 0 ? PG_HDR_FIRST : PG_HDR, content = PG_PAGE - hdr;
            const uint32_t l = content < left ? content : left;

            /* and it must be the chain the pages themselves describe */
            if (ld4(pg + PG_OFF_FIRST) != blkno[c0] || (k + 1 < np && ld4(pg + PG_OFF_NEXT) != blkno[c0 + k + 1]))
                st = PG_ST_CHAIN;
            /* whole vectors: the last payload is copied up to the next multiple of 16 (still inside its page) */
            for (uint32_t v = tid; v < (l + 15u) / 16u; v += nthr)
                st16(dst + 16u * v, ld16(pg + hdr + 16u * v));
            dst += l;
            left -= l;
        }
    }
    if (tid == 0)
    {
        *src_off = (uint64_t) c0 * PG_PAGE;
        *src_size = st == PG_ST_OK ? csize : 0u;
        *dec_method = st == PG_ST_OK ? (int32_t) method : PG_METHOD_SKIP;
        *hdr_method = (int32_t) method;
        *chain_status = st;
    }
}

/*
 * Split, one CTA, one cryo block: the page images cryo_preserve writes (pg_cryogen.c:761-805) for the
 * `size` compressed bytes at comp, into out[0, npages * PG_PAGE).  blkno[0, npages): the block numbers the
 * pages will have (the caller reserved cap_pages of them).  Returns the number of pages, 0 when they do
 * not fit cap_pages.  comp and out are 16-byte aligned.
 */
CRYO_DEV uint32_t pg_split_block(const uint8_t *comp, uint32_t size, uint32_t method, uint32_t xid, const uint32_t *blkno,
                                 uint32_t cap_pages, uint8_t *out, uint32_t tid, uint32_t nthr)
{
    const uint32_t np = pg_pages_needed(size);
    uint32_t left = size;
    const uint8_t *p = comp;

    if (np > cap_pages || np > 0xFFFFu)
        return 0;
    for (uint32_t k = 0; k < np; k++)
    {
        uint8_t       *pg = out + (size_t) k * PG_PAGE;
        const uint32_t hdr = k == 0 ? PG_HDR_FIRST : PG_HDR, content = PG_PAGE - hdr;
        const uint32_t l = content < left ? content : left;

        /* header: three vectors (the third only on the first page) */
        if (tid < hdr / 16u)
        {
            uint4 v = make_uint4(0u, 0u, 0u, 0u);

            if (tid == 0)
                v.w = (hdr + l) | (PG_PAGE << 16);                      /* pd_lower | pd_upper */
            else if (tid == 1)
            {
                v.x = PG_PAGE & 0xFFFFu;                                /* pd_special | pd_pagesize_version (left 0) */
                v.z = blkno[0];                                         /* first */
                v.w = k + 1 < np ? blkno[k + 1] : PG_INVALID;           /* next */
            }
            else
            {
                v.x = xid;
                v.y = method;
                v.z = size;
                v.w = np;                                               /* npages (uint16) + padding */
            }
            st16(pg + 16u * tid, v);
        }
        /* payload, then zeros to the end of the page */
        const uint32_t full = l / 16u, nvec = content / 16u;

        for (uint32_t v = tid; v < nvec; v += nthr)
        {
            uint4 x = make_uint4(0u, 0u, 0u, 0u);

            if (v < full)
                x = ld16(p + 16u * v);
            else if (v == full && (l & 15u))
            {
                uint32_t w[4] = {0u, 0u, 0u, 0u};

                for (uint32_t i = 0; i < (l & 15u); i++)
                    w[i >> 2] |= (uint32_t) p[16u * v + i] << (8u * (i & 3u));
                x = make_uint4(w[0], w[1], w[2], w[3]);
            }
            st16(pg + hdr + 16u * v, x);
        }
        p += l;
        left -= l;
    }
    return np;
}

/*
 * Tuple-level walk of one decoded cryo block by one warp: the items a sequential scan
 * returns (cryo_getnextslot, pg_cryogen.c:293: item cur_item while cur_item * sizeof(CryoItemId) <
 * hdr->lower; cryo_storage_fetch, storage.c:55-68: CryoItemId {off, len} at data + (pos - 1)).  The
 * block is in HBM (it was just decoded: L2): a pushed-down count / size aggregate then returns 16 bytes
 * per block over the bus instead of the block.  valid: every item lies inside [upper, block_size).
 */
CRYO_DEV void pg_tuple_stats(const uint8_t *block, uint32_t block_size, uint32_t *ntuples, unsigned long long *tuple_bytes,
                             int32_t *valid, uint32_t lane)
{
    const uint32_t lower = ld4(block), upper = ld4(block + 4);
    uint32_t n = lower ? (lower - 1u) / 8u : 0u;       /* largest cur_item with cur_item * 8 < lower */
    unsigned long long bytes = 0;
    bool ok = lower >= 8u && lower <= upper && upper <= block_size;

    if ((unsigned long long) n * 8u + 8u > block_size)
        n = block_size / 8u - 1u;
    for (uint32_t cur = 1u + lane; cur <= n; cur += 32u)
    {
        const uint2 it = *reinterpret_cast<const uint2 *>(block + 8u * cur);

        bytes += it.y;
        if (it.x < upper || it.x > block_size || it.y > block_size - it.x)
            ok = false;
    }
#pragma unroll
    for (uint32_t d = 16; d; d >>= 1)
        bytes += __shfl_xor_sync(CRYO_FULL, bytes, d);
    ok = __all_sync(CRYO_FULL, ok);
    if (lane == 0)
    {
        *ntuples = n;
        *tuple_bytes = bytes;
        *valid = ok ? 1 : 0;
    }
}
