/*
 * tests/shim_harness.c -- TEST INFRASTRUCTURE: calls the drop-in libcryo_compression.so the way
 * a PostgreSQL backend would (reference call sites pg_cryogen.c:726 and cache.c:178), with the
 * header shim's elog(ERROR) longjmp armed so that the error paths of compression.c:73-74, :137
 * and :157 come back to the test as return codes instead of aborting the process.
 */
#include "postgres.h"
#include "compression.h"

extern void cryo_compression_shutdown(void);
extern int  cryo_gpu_device_guc;

/* rc 0: ok; 1: elog(ERROR) was raised, msg holds the text */
int
shim_define_gucs(int *method, int *accel, int *level)
{
    cryo_define_compression_gucs();
    *method = compression_method_guc;
    *accel = lz4_acceleration_guc;
    *level = zstd_compression_level_guc;
    return 0;
}

void
shim_set_gucs(int method, int accel, int level)
{
    compression_method_guc = method;
    lz4_acceleration_guc = accel;
    zstd_compression_level_guc = level;
}

int
shim_compress(int method, const char *data, char *out, size_t cap, size_t *size, char *msg, size_t msgcap)
{
    sigjmp_buf  jb;
    char       *c;
    Size        sz = 0;

    pg_shim_error_jmp = &jb;
    if (sigsetjmp(jb, 0))
    {
        pg_shim_error_jmp = NULL;
        snprintf(msg, msgcap, "%s", pg_shim_last_error);
        return 1;
    }
    c = cryo_compress((CompressionMethod) method, data, &sz);
    pg_shim_error_jmp = NULL;
    if (c == NULL)
        return 2;
    if (sz > cap)
    {
        pfree(c);
        return 3;
    }
    memcpy(out, c, sz);
    pfree(c);               /* the caller's side of the ownership contract: pg_cryogen.c:826 */
    *size = sz;
    return 0;
}

/* rc 0: returned true; -1: returned false; 1: elog(ERROR) */
int
shim_decompress(int method, const char *compressed, size_t csize, char *out, char *msg, size_t msgcap)
{
    sigjmp_buf jb;
    bool       ok;

    pg_shim_error_jmp = &jb;
    if (sigsetjmp(jb, 0))
    {
        pg_shim_error_jmp = NULL;
        snprintf(msg, msgcap, "%s", pg_shim_last_error);
        return 1;
    }
    ok = cryo_decompress((CompressionMethod) method, compressed, csize, out);
    pg_shim_error_jmp = NULL;
    return ok ? 0 : -1;
}

void
shim_shutdown(void)
{
    cryo_compression_shutdown();
}
