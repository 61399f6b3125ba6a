/*
 * compression.h -- the codec boundary of pg_cryogen, as kept by the B200 drop-in.
 *
 * Same names, argument meaning and error behaviour as the reference's
 * compression.h:7-24, so that pg_cryogen.c:726 (cryo_preserve) and cache.c:178
 * (cryo_read_decompress) link against host/compression.c unchanged.  When this
 * file is dropped into the reference tree the reference's own compression.h can
 * be used instead; the two declare the same interface.
 */
#ifndef __COMPRESSION_H__
#define __COMPRESSION_H__

#include "postgres.h"

/* stored on disk as a 4-byte int in CryoFirstPageHeader (storage.h:64) */
typedef enum
{
    COMP_LZ4 = 0,
    COMP_ZSTD
} CompressionMethod;

/* GUCs: pg_cryogen.compression_method / lz4_acceleration / zstd_compression_level */
extern int compression_method_guc;
extern int lz4_acceleration_guc;
extern int zstd_compression_level_guc;

/* compress one CRYO_BLCKSZ block; returns a palloc'd buffer, elog(ERROR) on failure */
extern char *cryo_compress(CompressionMethod method, const char *data, Size *compressed_size);
/* decompress into `out` (capacity CRYO_BLCKSZ); false on a malformed stream */
extern bool cryo_decompress(CompressionMethod method, const char *compressed,
                            Size compressed_size, char *out);
extern void cryo_define_compression_gucs(void);

#endif /* __COMPRESSION_H__ */
