/*
 * oracle/cryo_pages.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of how pg_cryogen lays a compressed cryo block out over PostgreSQL pages
 * and reads it back, for the parity tests of cryogpu_*_pages_* (SURVEY.md 8 f-1, a10):
 *
 *   split    cryo_pages_needed / cryo_preserve, /root/reference/pg_cryogen.c:689-704, :761-805
 *   gather   cryo_read_decompress,              /root/reference/cache.c:100-176
 *   layout   PageHeaderClone, CryoPageHeader, CryoFirstPageHeader, /root/reference/storage.h:26-67
 *
 * Those functions live inside the table access method (buffer manager, WAL, locks) and cannot be
 * compiled outside a PostgreSQL backend, so they are restated here over a malloc'ed "relation"
 * (an array of 8 KiB pages indexed by block number).  Parity pinning: the structure sizes and
 * field offsets below are checked against the reference's own storage.h compiled in this
 * container (oref_page_layout in oracle/ref_driver.c -> oracle/_ref), tests/test_pages.py; the
 * reference holds no golden vectors for page images.
 */
#include <stdint.h>
#include <string.h>

#define PG_BLCKSZ           8192u
#define PG_INVALID_BLOCK    0xFFFFFFFFu

/* storage.h:26-67, as laid out by the compiler on x86-64 / aarch64 (little endian) */
#define OFF_PD_LOWER        12u     /* PageHeaderClone.pd_lower   (uint16) */
#define OFF_PD_UPPER        14u     /* PageHeaderClone.pd_upper   (uint16) */
#define OFF_PD_SPECIAL      16u     /* PageHeaderClone.pd_special (uint16) */
#define OFF_FIRST           24u     /* CryoPageHeader.first       (BlockNumber) */
#define OFF_NEXT            28u     /* CryoPageHeader.next        (BlockNumber) */
#define SZ_PAGE_HEADER      32u     /* sizeof(CryoPageHeader) */
#define OFF_CREATED_XID     32u     /* CryoFirstPageHeader.created_xid        (TransactionId) */
#define OFF_METHOD          36u     /* CryoFirstPageHeader.compression_method (enum, 4 bytes) */
#define OFF_COMP_SIZE       40u     /* CryoFirstPageHeader.compressed_size    (uint32) */
#define OFF_NPAGES          44u     /* CryoFirstPageHeader.npages             (uint16) */
#define SZ_FIRST_HEADER     48u     /* sizeof(CryoFirstPageHeader) */

/* cache.c error codes (cache.h) */
#define CRYO_ERR_SUCCESS                0
#define CRYO_ERR_WRONG_STARTING_BLOCK   2
#define CRYO_ERR_EMPTY_BLOCK            3

static void put16(uint8_t *p, uint32_t v) { p[0] = (uint8_t) v; p[1] = (uint8_t) (v >> 8); }
static void put32(uint8_t *p, uint32_t v) { put16(p, v & 0xFFFFu); put16(p + 2, v >> 16); }
static uint32_t get16(const uint8_t *p) { return p[0] | ((uint32_t) p[1] << 8); }
static uint32_t get32(const uint8_t *p) { return get16(p) | (get16(p + 2) << 16); }

void
cryo_oracle_page_layout(uint32_t out[12])
{
    out[0] = SZ_PAGE_HEADER; out[1] = SZ_FIRST_HEADER; out[2] = OFF_FIRST; out[3] = OFF_NEXT;
    out[4] = OFF_CREATED_XID; out[5] = OFF_METHOD; out[6] = OFF_COMP_SIZE; out[7] = OFF_NPAGES;
    out[8] = OFF_PD_LOWER; out[9] = OFF_PD_UPPER; out[10] = OFF_PD_SPECIAL; out[11] = PG_BLCKSZ;
}

/* pg_cryogen.c:692-704.  (The reference returns uint8, which wraps for blocks over 255 pages; a
 * 1 MiB cryo block needs at most 130.) */
uint32_t
cryo_oracle_pages_needed(uint64_t size)
{
    uint32_t pages = 1;
    const uint64_t page_sz = PG_BLCKSZ - SZ_PAGE_HEADER, first_sz = PG_BLCKSZ - SZ_FIRST_HEADER;

    if (size > first_sz)
        pages += (uint32_t) ((size - first_sz + page_sz - 1) / page_sz);
    return pages;
}

/*
 * pg_cryogen.c:761-805: the compressed bytes of one cryo block into npages fresh (zeroed) pages of the
 * relation, at block numbers blkno[0..npages).  rel: the relation, nrel pages.  Returns npages, or 0
 * when a block number is outside the relation.
 */
uint32_t
cryo_oracle_pages_split(uint8_t *rel, uint32_t nrel, const uint32_t *blkno, const uint8_t *comp, uint32_t size,
                        uint32_t method, uint32_t xid)
{
    const uint32_t npages = cryo_oracle_pages_needed(size);
    uint32_t left = size;
    const uint8_t *p = comp;

    for (uint32_t i = 0; i < npages; i++)
    {
        if (blkno[i] >= nrel)
            return 0;
        uint8_t *hdr = rel + (uint64_t) blkno[i] * PG_BLCKSZ;
        const uint32_t hdr_size = i == 0 ? SZ_FIRST_HEADER : SZ_PAGE_HEADER;    /* CryoPageHeaderSize: first == block */
        const uint32_t content = PG_BLCKSZ - hdr_size, take = content < left ? content : left;

        memset(hdr, 0, PG_BLCKSZ);                                  /* a page from ReadBuffer(P_NEW) */
        put32(hdr + OFF_FIRST, blkno[0]);
        put32(hdr + OFF_NEXT, i + 1 < npages ? blkno[i + 1] : PG_INVALID_BLOCK);
        if (i == 0)
        {
            put16(hdr + OFF_NPAGES, npages);
            put32(hdr + OFF_METHOD, method);
            put32(hdr + OFF_COMP_SIZE, size);
            put32(hdr + OFF_CREATED_XID, xid);
        }
        put16(hdr + OFF_PD_UPPER, PG_BLCKSZ & 0xFFFFu);             /* LocationIndex is uint16: 8192 fits */
        put16(hdr + OFF_PD_LOWER, hdr_size + take);
        put16(hdr + OFF_PD_SPECIAL, PG_BLCKSZ & 0xFFFFu);
        memcpy(hdr + hdr_size, p, take);
        left -= take;
        p += take;
    }
    return npages;
}

/*
 * cache.c:100-176: the compressed bytes of the cryo block that starts at `block`.  out: at least
 * compressed_size bytes (cap).  *got: bytes gathered (less than *size when the chain ends early: the
 * reference then hands the short buffer to cryo_decompress, which fails).  blocks[]: the chain.
 */
int
cryo_oracle_pages_gather(const uint8_t *rel, uint32_t nrel, uint32_t block, uint8_t *out, uint32_t cap,
                         uint32_t *method, uint32_t *size, uint32_t *got, uint32_t *blocks, uint32_t *nblocks)
{
    *nblocks = 0;
    *got = 0;
    if (block >= nrel)
        return CRYO_ERR_EMPTY_BLOCK;
    const uint8_t *page = rel + (uint64_t) block * PG_BLCKSZ;

    if (get16(page + OFF_PD_UPPER) == 0)                            /* PageIsNew */
        return CRYO_ERR_EMPTY_BLOCK;
    if (get32(page + OFF_FIRST) != block)
        return CRYO_ERR_WRONG_STARTING_BLOCK;
    uint32_t left = get32(page + OFF_COMP_SIZE), cur = block;
    uint8_t *p = out;

    *size = left;
    *method = get32(page + OFF_METHOD);
    if (left > cap)
        left = cap;
    blocks[(*nblocks)++] = block;
    for (;;)
    {
        const uint32_t hdr_size = get32(page + OFF_FIRST) == cur ? SZ_FIRST_HEADER : SZ_PAGE_HEADER;
        const uint32_t content = PG_BLCKSZ - hdr_size, l = content < left ? content : left;

        memcpy(p, page + hdr_size, l);
        p += l;
        left -= l;
        cur = get32(page + OFF_NEXT);
        if (left == 0)
            break;
        if (cur == PG_INVALID_BLOCK || cur >= nrel)
            break;
        page = rel + (uint64_t) cur * PG_BLCKSZ;
        blocks[(*nblocks)++] = cur;
    }
    *got = (uint32_t) (p - out);
    return CRYO_ERR_SUCCESS;
}

/*
 * Tuple-level walk of one decoded cryo block: what a sequential scan visits.
 * storage.h:69-86: CryoDataHeader {uint32 lower, upper; char data[]}, CryoItemId {uint32 off, len};
 * item positions are 1-based (storage.c:55-68); cryo_getnextslot returns item cur_item while
 * cur_item * sizeof(CryoItemId) < hdr->lower (pg_cryogen.c:293).  valid: every item lies inside
 * [upper, block_size) -- the reference only asserts this.
 */
void
cryo_oracle_block_tuple_stats(const uint8_t *block, uint32_t block_size, uint32_t *ntuples, uint64_t *tuple_bytes,
                              int32_t *valid)
{
    const uint32_t lower = get32(block), upper = get32(block + 4);
    uint32_t n = 0;
    uint64_t bytes = 0;
    int      ok = lower >= 8 && lower <= upper && upper <= block_size;

    for (uint32_t cur = 1; (uint64_t) cur * 8u < lower && (uint64_t) cur * 8u + 8u <= block_size; cur++)
    {
        const uint32_t off = get32(block + 8u * cur), len = get32(block + 8u * cur + 4);

        bytes += len;
        n++;
        if (off < upper || off > block_size || len > block_size - off)
            ok = 0;
    }
    *ntuples = n;
    *tuple_bytes = bytes;
    *valid = ok;
}
