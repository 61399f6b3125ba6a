/* pg_shim/zstd.h -- prototypes of the libzstd 1.5.5 entry points compression.c calls
 * (the image ships libzstd.so.1 without headers). */
#ifndef PG_SHIM_ZSTD_H
#define PG_SHIM_ZSTD_H
#include <stddef.h>
unsigned ZSTD_versionNumber(void);
size_t ZSTD_compressBound(size_t srcSize);
size_t ZSTD_compress(void *dst, size_t dstCapacity, const void *src, size_t srcSize, int compressionLevel);
size_t ZSTD_decompress(void *dst, size_t dstCapacity, const void *src, size_t compressedSize);
unsigned ZSTD_isError(size_t code);
const char *ZSTD_getErrorName(size_t code);
#endif
