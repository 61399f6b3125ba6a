/*
 * compression.c -- drop-in replacement for the reference's compression.c
 * (compression.c:1-159) that routes the block codec to libcryogpu.so on a B200.
 *
 * What stays exactly as in the reference: the exported symbols and their
 * signatures (compression.h:13-24), the three GUCs with their names, defaults
 * and ranges (compression.c:16-58), ownership (compress returns a palloc'd
 * buffer the caller pfree's, pg_cryogen.c:826; decompress fills the caller's
 * CRYO_BLCKSZ buffer, cache.c:46), and the error convention (compress raises
 * elog(ERROR), decompress returns false; compression.c:73-74, :85-86).
 *
 * What changes: no liblz4 / libzstd call.  Each call hands one block to
 * cryogpu_compress_host / cryogpu_decompress_host.  The GPU context is created
 * lazily on first use, i.e. inside the backend process after fork()
 * (pg_cryogen.c:169-176); nothing touches CUDA at load time.  All elog() calls
 * are made here, after the library has returned: no longjmp ever crosses a CUDA
 * runtime frame.  There is no CPU fallback: without a usable B200 the call
 * raises ERROR.
 */
#include "postgres.h"
#include "utils/guc.h"

#include "compression.h"
#include "cryogpu.h"

#ifndef CRYO_BLCKSZ
#define CRYO_BLCKSZ (1 << 20)       /* storage.h:18 */
#endif

static const struct config_enum_entry compression_method_options[] = {
    {"lz4", COMP_LZ4, false},
    {"zstd", COMP_ZSTD, false},
    {NULL, 0, false}
};

int compression_method_guc = COMP_ZSTD;
int lz4_acceleration_guc = 1;
int zstd_compression_level_guc = 1;

/* pg_cryogen.gpu_device: which CUDA device this backend uses */
int cryo_gpu_device_guc = 0;

static cryogpu_ctx *gpu_ctx = NULL;

void
cryo_define_compression_gucs(void)
{
    DefineCustomEnumVariable("pg_cryogen.compression_method",
                             "Possible values are lz4 and zstd.",
                             NULL, &compression_method_guc, COMP_ZSTD,
                             compression_method_options, PGC_USERSET, 0, NULL, NULL, NULL);
    DefineCustomIntVariable("pg_cryogen.lz4_acceleration", "Sets lz4 acceleration.",
                            NULL, &lz4_acceleration_guc, 1, 0, 50,
                            PGC_USERSET, 0, NULL, NULL, NULL);
    DefineCustomIntVariable("pg_cryogen.zstd_compression_level", "Sets zstd compression level.",
                            NULL, &zstd_compression_level_guc, 1, -5, 22,
                            PGC_USERSET, 0, NULL, NULL, NULL);
    DefineCustomIntVariable("pg_cryogen.gpu_device", "CUDA device used by this backend.",
                            NULL, &cryo_gpu_device_guc, 0, 0, 63,
                            PGC_USERSET, 0, NULL, NULL, NULL);
}

static cryogpu_ctx *
gpu(void)
{
    if (gpu_ctx == NULL)
    {
        int rc = cryogpu_init(cryo_gpu_device_guc, &gpu_ctx);

        if (rc != CRYOGPU_OK)
        {
            gpu_ctx = NULL;
            elog(ERROR, "pg_cryogen: GPU codec unavailable: %s", cryogpu_last_error());
        }
    }
    return gpu_ctx;
}

/* release the context (tests; a backend simply exits) */
void
cryo_compression_shutdown(void)
{
    if (gpu_ctx)
        cryogpu_shutdown(gpu_ctx);
    gpu_ctx = NULL;
}

char *
cryo_compress(CompressionMethod method, const char *data, Size *compressed_size)
{
    cryogpu_ctx *ctx;
    uint64_t     estimate;
    char        *compressed;
    const void  *src[1];
    void        *dst[1];
    uint32_t     out_size = 0;
    int32_t      status = -1;
    int          level, rc;

    switch (method)
    {
        case COMP_LZ4:
            level = lz4_acceleration_guc;
            break;
        case COMP_ZSTD:
            level = zstd_compression_level_guc;
            break;
        default:
            elog(ERROR, "pg_cryogen: unknown compression method");
            return NULL;
    }
    ctx = gpu();
    estimate = cryogpu_compress_bound((int) method, CRYO_BLCKSZ);
    compressed = palloc(estimate);
    src[0] = data;
    dst[0] = compressed;
    rc = cryogpu_compress_host(ctx, 1, (int) method, level, src, CRYO_BLCKSZ, dst,
                               (uint32_t) estimate, &out_size, &status);
    if (rc != CRYOGPU_OK || status != CRYOGPU_ST_OK || out_size == 0)
    {
        pfree(compressed);
        elog(ERROR, "pg_cryogen: compression failed");
        return NULL;
    }
    *compressed_size = out_size;
    return compressed;
}

/*
 * Decompress and store result in `out`
 */
bool
cryo_decompress(CompressionMethod method, const char *compressed, Size compressed_size, char *out)
{
    const void *src[1];
    void       *dst[1];
    int32_t     m = (int32_t) method, status = -1;
    uint32_t    csize = (uint32_t) compressed_size, out_size = 0;
    int         rc;

    if (method != COMP_LZ4 && method != COMP_ZSTD)
    {
        elog(ERROR, "pg_cryogen: unknown compression method");
        return false;
    }
    src[0] = compressed;
    dst[0] = out;
    rc = cryogpu_decompress_host(gpu(), 1, &m, src, &csize, dst, CRYO_BLCKSZ, &out_size, &status);
    if (rc != CRYOGPU_OK)
        elog(ERROR, "pg_cryogen: GPU codec failed: %s", cryogpu_last_error());
    /* like the reference, a short output is not an error in production builds
     * (compression.c:88, :120 only Assert the size) */
    return status == CRYOGPU_ST_OK;
}
