/* pg_shim/lz4.h -- prototypes of the liblz4 1.9.4 entry points compression.c calls
 * (the image ships liblz4.so.1 without headers). */
#ifndef PG_SHIM_LZ4_H
#define PG_SHIM_LZ4_H
int LZ4_versionNumber(void);
int LZ4_compressBound(int inputSize);
int LZ4_compress_fast(const char *src, char *dst, int srcSize, int dstCapacity, int acceleration);
int LZ4_decompress_safe(const char *src, char *dst, int compressedSize, int dstCapacity);
#endif
