/*
 * oracle/cryo_oracle.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of the two stream formats pg_cryogen stores on disk
 * (reference compression.c:61-123 delegates all of the arithmetic to the
 * third-party liblz4 / libzstd, which are not under /root/reference):
 *
 *   - the raw LZ4 block format, as decoded by LZ4_decompress_safe
 *     (called at compression.c:84); pinned library: liblz4 1.9.4 (image .so)
 *   - the zstd frame format, RFC 8878, as decoded by ZSTD_decompress
 *     (called at compression.c:116); pinned library: libzstd 1.5.5 (image .so)
 *
 * Parity pinning: the reference holds no golden vectors for this path
 * (SURVEY.md 8(c)); the port is pinned against outputs of the reference itself
 * run in this container (oracle/_ref/libcryoref.so) in tests/test_oracle.py, and
 * against the committed fixtures in tests/golden/.
 */
#ifndef CRYO_ORACLE_H
#define CRYO_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#define CRYO_ORACLE_OK 0
#define CRYO_ORACLE_ERR_INPUT      (-1)  /* truncated / overrun / not consumed exactly */
#define CRYO_ORACLE_ERR_OUTPUT     (-2)  /* output would exceed capacity */
#define CRYO_ORACLE_ERR_OFFSET     (-3)  /* match offset 0 or before start of output */
#define CRYO_ORACLE_ERR_FORMAT     (-4)  /* reserved value / bad magic / bad table */
#define CRYO_ORACLE_ERR_SIZE       (-5)  /* frame content size mismatch */

/* Counters filled in while decoding (what kinds of things the stream held). */
typedef struct
{
    uint32_t sequences;
    uint32_t literal_bytes;
    uint32_t match_bytes;
    uint32_t longest_match;
    uint32_t overlap_match_bytes;   /* bytes of matches with offset < length */
    uint32_t max_literal_run;
    uint32_t max_offset;
    /* zstd only */
    uint32_t frames;
    uint32_t window_size;
    uint32_t single_segment;
    uint32_t blocks_raw, blocks_rle, blocks_compressed;
    uint32_t lit_raw, lit_rle, lit_huf1, lit_huf4, lit_treeless1, lit_treeless4;
    uint32_t huf_direct_weights, huf_fse_weights;
    uint32_t mode_predef, mode_rle, mode_fse, mode_repeat;   /* summed over LL/OF/ML */
    uint32_t rep_offsets;           /* sequences that used a repeat offset */
} cryo_oracle_stats;

/* Sequence trace callback (zstd and LZ4): lit_len, match_len, offset (resolved). */
typedef void (*cryo_oracle_seq_cb)(void *ctx, uint32_t lit_len, uint32_t match_len, uint32_t offset);

/*
 * LZ4 block decode with LZ4_decompress_safe's acceptance rules.
 * Returns the number of bytes written (>= 0) or a negative CRYO_ORACLE_ERR_*.
 */
long cryo_oracle_lz4_decode(const uint8_t *src, size_t src_size, uint8_t *dst, size_t dst_cap,
                            cryo_oracle_stats *st, cryo_oracle_seq_cb cb, void *cb_ctx);

/*
 * zstd frame(s) decode (RFC 8878; no dictionary).  Accepts concatenated and
 * skippable frames like ZSTD_decompress.  Returns bytes written or negative.
 */
long cryo_oracle_zstd_decode(const uint8_t *src, size_t src_size, uint8_t *dst, size_t dst_cap,
                             cryo_oracle_stats *st, cryo_oracle_seq_cb cb, void *cb_ctx);

/* XXH64 (zstd content checksum = low 32 bits of XXH64 seed 0). */
uint64_t cryo_oracle_xxh64(const uint8_t *p, size_t n, uint64_t seed);

#endif
