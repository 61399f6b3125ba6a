/*
 * oracle/ref_driver.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Batch + pthread driver around the UNMODIFIED reference sources
 * /root/reference/compression.c and /root/reference/storage.c, which the
 * Makefile compiles where they lie against pg_shim/ and links with the
 * image's liblz4.so.1 (1.9.4) / libzstd.so.1 (1.5.5) into oracle/_ref/libcryoref.so.
 *
 * Every codec call below goes through the reference's own entry points
 * cryo_compress (compression.c:125-139) and cryo_decompress (compression.c:144-159),
 * and block building goes through cryo_init_page / cryo_storage_insert
 * (storage.c:15-50).  Nothing here re-implements the reference.
 */
#define _GNU_SOURCE
#include "storage.h"      /* reference header: CRYO_BLCKSZ, CryoDataHeader */
#include "compression.h"  /* reference header: cryo_compress / cryo_decompress */

#include <pthread.h>
#include <time.h>

extern int LZ4_versionNumber(void);
extern unsigned ZSTD_versionNumber(void);
extern int LZ4_compressBound(int);
extern size_t ZSTD_compressBound(size_t);

static double
now_sec(void)
{
    struct timespec ts;

    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + ts.tv_nsec * 1e-9;
}

void
oref_versions(int *lz4, int *zstd)
{
    *lz4 = LZ4_versionNumber();
    *zstd = (int) ZSTD_versionNumber();
}

uint64_t
oref_block_size(void)
{
    return CRYO_BLCKSZ;
}

/* sizes and field offsets of the page headers, from the reference's own storage.h (pins oracle/cryo_pages.c) */
void
oref_page_layout(uint32_t out[12])
{
    out[0] = sizeof(CryoPageHeader);
    out[1] = sizeof(CryoFirstPageHeader);
    out[2] = offsetof(CryoPageHeader, first);
    out[3] = offsetof(CryoPageHeader, next);
    out[4] = offsetof(CryoFirstPageHeader, created_xid);
    out[5] = offsetof(CryoFirstPageHeader, compression_method);
    out[6] = offsetof(CryoFirstPageHeader, compressed_size);
    out[7] = offsetof(CryoFirstPageHeader, npages);
    out[8] = offsetof(PageHeaderClone, pd_lower);
    out[9] = offsetof(PageHeaderClone, pd_upper);
    out[10] = offsetof(PageHeaderClone, pd_special);
    out[11] = BLCKSZ;
}

/*
 * The item walk of a sequential scan over one decoded block: the loop condition of cryo_getnextslot
 * (pg_cryogen.c:293) around the reference's own cryo_storage_fetch (storage.c:55-68).  Pins
 * cryo_oracle_block_tuple_stats and the GPU's k_tuple_stats.
 */
void
oref_block_walk(char *block, uint32_t *ntuples, uint64_t *tuple_bytes)
{
    CryoDataHeader *hdr = (CryoDataHeader *) block;
    uint32_t    cur_item = 1, n = 0;
    uint64_t    bytes = 0;

    while (cur_item * sizeof(CryoItemId) < hdr->lower)
    {
        HeapTupleData t;

        cryo_storage_fetch(hdr, cur_item, &t);
        bytes += t.t_len;
        n++;
        cur_item++;
    }
    *ntuples = n;
    *tuple_bytes = bytes;
}

/* bound the reference allocates: compression.c:67 / compression.c:99 */
uint64_t
oref_compress_bound(int method)
{
    return method == COMP_LZ4 ? (uint64_t) LZ4_compressBound(CRYO_BLCKSZ)
                              : (uint64_t) ZSTD_compressBound(CRYO_BLCKSZ);
}

void
oref_define_gucs(int *method, int *accel, int *level)
{
    cryo_define_compression_gucs();
    *method = compression_method_guc;
    *accel = lz4_acceleration_guc;
    *level = zstd_compression_level_guc;
}

/* ---- block building through the reference's storage.c ---- */

void
oref_init_page(char *block)
{
    cryo_init_page((CryoDataHeader *) block);
}

int
oref_storage_insert(char *block, const char *tuple, uint32_t len)
{
    HeapTupleData t;

    memset(&t, 0, sizeof(t));
    t.t_len = len;
    t.t_data = (HeapTupleHeader) tuple;
    return cryo_storage_insert((CryoDataHeader *) block, &t);
}

/* returns length, copies tuple bytes to out (capacity cap) */
int
oref_storage_fetch(char *block, int pos, char *out, uint32_t cap)
{
    HeapTupleData t;

    cryo_storage_fetch((CryoDataHeader *) block, pos, &t);
    if (t.t_len <= cap)
        memcpy(out, t.t_data, t.t_len);
    return (int) t.t_len;
}

/* ---- batched codec calls ---- */

typedef struct
{
    int         method;         /* compress: method for all blocks */
    const int  *methods;        /* decompress: per block */
    const char *src;
    const uint64_t *src_off;
    const uint32_t *src_size;
    char       *dst;
    uint64_t    dst_stride;
    uint32_t   *dst_size;
    uint8_t    *ok;
    size_t      begin, end;
    int         reps;
} Job;

static void *
compress_worker(void *arg)
{
    Job *j = (Job *) arg;

    for (int r = 0; r < j->reps; r++)
        for (size_t i = j->begin; i < j->end; i++)
        {
            Size sz = 0;
            char *c = cryo_compress((CompressionMethod) j->method,
                                    j->src + i * (uint64_t) CRYO_BLCKSZ, &sz);

            if (j->dst)
                memcpy(j->dst + i * j->dst_stride, c, sz);
            j->dst_size[i] = (uint32_t) sz;
            pfree(c);
        }
    return NULL;
}

static void *
decompress_worker(void *arg)
{
    Job *j = (Job *) arg;

    for (int r = 0; r < j->reps; r++)
        for (size_t i = j->begin; i < j->end; i++)
        {
            bool ok = cryo_decompress((CompressionMethod) j->methods[i],
                                      j->src + j->src_off[i], j->src_size[i],
                                      j->dst + i * j->dst_stride);

            j->ok[i] = ok ? 1 : 0;
        }
    return NULL;
}

static double
run_jobs(Job *proto, size_t n, int nthreads, void *(*fn)(void *))
{
    pthread_t  *th;
    Job        *jobs;
    double      t0;

    if (nthreads < 1)
        nthreads = 1;
    if ((size_t) nthreads > n && n > 0)
        nthreads = (int) n;
    th = calloc(nthreads, sizeof(*th));
    jobs = calloc(nthreads, sizeof(*jobs));
    t0 = now_sec();
    for (int t = 0; t < nthreads; t++)
    {
        jobs[t] = *proto;
        jobs[t].begin = n * t / nthreads;
        jobs[t].end = n * (t + 1) / nthreads;
        if (nthreads == 1)
            fn(&jobs[t]);
        else
            pthread_create(&th[t], NULL, fn, &jobs[t]);
    }
    if (nthreads > 1)
        for (int t = 0; t < nthreads; t++)
            pthread_join(th[t], NULL);
    t0 = now_sec() - t0;
    free(th);
    free(jobs);
    return t0;
}

/*
 * Compress n contiguous 1 MiB blocks at src.  level_or_accel is stored into
 * the reference's GUC variable that compression.c:72 / :104 reads.  dst may be
 * NULL (timing only).  Returns elapsed seconds.
 */
double
oref_compress_batch(int method, int level_or_accel, const char *src, size_t n,
                    char *dst, uint64_t dst_stride, uint32_t *dst_size,
                    int nthreads, int reps)
{
    Job j;

    memset(&j, 0, sizeof(j));
    if (method == COMP_LZ4)
        lz4_acceleration_guc = level_or_accel;
    else
        zstd_compression_level_guc = level_or_accel;
    j.method = method;
    j.src = src;
    j.dst = dst;
    j.dst_stride = dst_stride;
    j.dst_size = dst_size;
    j.reps = reps < 1 ? 1 : reps;
    return run_jobs(&j, n, nthreads, compress_worker);
}

/*
 * Decompress n blocks; block i is src[src_off[i] .. +src_size[i]) with method
 * methods[i]; output block i goes to dst + i*dst_stride (capacity CRYO_BLCKSZ).
 * ok[i] receives cryo_decompress's bool.  Returns elapsed seconds.
 */
double
oref_decompress_batch(const int *methods, const char *src, const uint64_t *src_off,
                      const uint32_t *src_size, size_t n, char *dst,
                      uint64_t dst_stride, uint8_t *ok, int nthreads, int reps)
{
    Job j;

    memset(&j, 0, sizeof(j));
    j.methods = methods;
    j.src = src;
    j.src_off = src_off;
    j.src_size = src_size;
    j.dst = dst;
    j.dst_stride = dst_stride;
    j.ok = ok;
    j.reps = reps < 1 ? 1 : reps;
    return run_jobs(&j, n, nthreads, decompress_worker);
}
