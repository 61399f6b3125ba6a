/*
 * cryogpu.cu -- C ABI of libcryogpu.so (include/cryogpu.h) and the sm_100a kernel
 * entry points.  Host side: lazy per-device context, pinned staging, chunked
 * H2D / kernel / D2H pipelines for the *_host calls, block-range sharding over
 * several GPUs.  No CPU codec lives here: every byte of LZ4 / zstd work is done
 * by the kernels in lz4_decode.cuh, zstd_decode.cuh, lz4_encode.cuh and
 * zstd_encode.cuh.
 */
#include "../../include/cryogpu.h"

#include "lz4_decode.cuh"
#include "lz4_decode_w.cuh"
#include "zstd_decode.cuh"
#include "zstd_decode_w.cuh"
#include "lz4_encode.cuh"
#include "zstd_encode.cuh"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

/* ------------------------------------------------------------------ errors */

static thread_local char g_err[512] = "";

static int
fail(int code, const char *fmt, ...)
{
    va_list ap;

    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                         \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess)                                                           \
            return fail(CRYOGPU_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                             \
    } while (0)

/* ----------------------------------------------------------------- kernels */

__global__ void __launch_bounds__(LZ4D_THREADS)
k_lz4_decode(const int32_t *methods, const uint8_t *src, const uint64_t *src_off,
             const uint32_t *src_size, uint8_t *dst, uint64_t dst_stride, uint32_t cap,
             uint32_t *out_size, int32_t *status)
{
    const uint32_t b = blockIdx.x;

    if (methods[b] != CRYOGPU_LZ4)
        return;
    lz4_decode_block(src + src_off[b], src_size[b], dst + b * dst_stride, cap, out_size + b,
                     status + b);
}

/* throughput path: one warp per block, LZ4W_WARPS blocks per CTA */
__global__ void __launch_bounds__(LZ4W_THREADS)
k_lz4_decode_w(const int32_t *methods, const uint8_t *src, const uint64_t *src_off,
               const uint32_t *src_size, uint8_t *dst, uint64_t dst_stride, uint32_t cap,
               uint32_t *out_size, int32_t *status, uint32_t n)
{
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t b = blockIdx.x * LZ4W_WARPS + warp;

    if (b >= n || methods[b] != CRYOGPU_LZ4)
        return;
    lz4w_decode_block(src + src_off[b], src_size[b], dst + b * dst_stride, cap, out_size + b,
                      status + b, CRYO_SMEM_BASE() + warp * LZ4W_PER_WARP, lane);
}

__global__ void __launch_bounds__(ZSTDD_THREADS)
k_zstd_decode(const int32_t *methods, const uint8_t *src, const uint64_t *src_off,
              const uint32_t *src_size, uint8_t *dst, uint64_t dst_stride, uint32_t cap,
              uint32_t *out_size, int32_t *status, uint8_t *scratch, uint64_t scratch_stride)
{
    const uint32_t b = blockIdx.x;

    if (methods[b] != CRYOGPU_ZSTD)
        return;
    zstd_decode_frame(src + src_off[b], src_size[b], dst + b * dst_stride, cap, out_size + b,
                      status + b, scratch + b * scratch_stride);
}

/* throughput path: one warp per frame, ZSW_WARPS frames per CTA */
__global__ void __launch_bounds__(ZSW_THREADS, ZSW_CTAS_PER_SM)
k_zstd_decode_w(const int32_t *methods, const uint8_t *src, const uint64_t *src_off,
                const uint32_t *src_size, uint8_t *dst, uint64_t dst_stride, uint32_t cap,
                uint32_t *out_size, int32_t *status, uint8_t *scratch, uint64_t scratch_stride,
                const uint32_t *predef, uint32_t n)
{
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t b = blockIdx.x * ZSW_WARPS + warp;

    if (b >= n || methods[b] != CRYOGPU_ZSTD)
        return;
    zstdw_decode_frame(src + src_off[b], src_size[b], dst + b * dst_stride, cap, out_size + b,
                       status + b, scratch + b * scratch_stride, predef,
                       CRYO_SMEM_BASE() + warp * ZSW_PER_WARP, lane);
}

/* the three predefined FSE tables of RFC 8878 3.1.1.3.2.2, built once per context */
__global__ void
k_zstd_build_predef(uint32_t *predef)
{
    __shared__ __align__(16) uint8_t sm[2048];

    zsw_build_predef(predef, sm, threadIdx.x);
}

__global__ void
k_flag_unknown_methods(const int32_t *methods, size_t n, uint32_t *out_size, int32_t *status)
{
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;

    if (i < n && methods[i] != CRYOGPU_LZ4 && methods[i] != CRYOGPU_ZSTD)
    {
        status[i] = CRYOGPU_ST_METHOD;      /* compression.c:157 "unknown compression method" */
        out_size[i] = 0;
    }
}

__global__ void __launch_bounds__(LZ4E_THREADS)
k_lz4_encode(const uint8_t *src, uint64_t src_stride, uint32_t block_size, uint8_t *dst,
             uint64_t dst_stride, uint32_t dst_cap, int accel, uint32_t *dst_size,
             int32_t *status, uint8_t *scratch, uint64_t scratch_stride)
{
    const uint32_t b = blockIdx.x;

    lz4_encode_block(src + b * src_stride, block_size, dst + b * dst_stride, dst_cap, accel,
                     dst_size + b, status + b, scratch + b * scratch_stride);
}

/* persistent: one CTA per SM walks the batch; scratch is per CTA, not per block */
__global__ void __launch_bounds__(ZSTDE_THREADS)
k_zstd_encode(const uint8_t *src, uint64_t src_stride, uint32_t block_size, uint8_t *dst,
              uint64_t dst_stride, uint32_t dst_cap, int level, uint32_t *dst_size,
              int32_t *status, uint8_t *scratch, uint64_t scratch_stride, uint32_t n)
{
    for (uint32_t b = blockIdx.x; b < n; b += gridDim.x)
        zstd_encode_frame(src + b * src_stride, block_size, dst + b * dst_stride, dst_cap, level,
                          dst_size + b, status + b, scratch + blockIdx.x * scratch_stride);
}

/* which LZ4 decode kernel: CRYOGPU_LZ4_KERNEL=cta selects the one-CTA-per-block variant */
static bool
lz4_use_cta_kernel()
{
    static int v = -1;

    if (v < 0)
    {
        const char *e = getenv("CRYOGPU_LZ4_KERNEL");

        v = (e && strcmp(e, "cta") == 0) ? 1 : 0;
    }
    return v == 1;
}

static void
launch_lz4_decode(cudaStream_t st, size_t n, const int32_t *methods, const uint8_t *src,
                  const uint64_t *src_off, const uint32_t *src_size, uint8_t *dst,
                  uint64_t dst_stride, uint32_t cap, uint32_t *out_size, int32_t *status)
{
    if (lz4_use_cta_kernel())
        k_lz4_decode<<<(unsigned) n, LZ4D_THREADS, LZ4D_SMEM, st>>>(methods, src, src_off, src_size,
                                                                   dst, dst_stride, cap, out_size,
                                                                   status);
    else
        k_lz4_decode_w<<<(unsigned) ((n + LZ4W_WARPS - 1) / LZ4W_WARPS), LZ4W_THREADS, LZ4W_SMEM, st>>>(
            methods, src, src_off, src_size, dst, dst_stride, cap, out_size, status, (uint32_t) n);
}

static bool
zstd_use_cta_kernel()
{
    static int v = -1;

    if (v < 0)
    {
        const char *e = getenv("CRYOGPU_ZSTD_KERNEL");

        v = (e && strcmp(e, "cta") == 0) ? 1 : 0;
    }
    return v == 1;
}

static void
launch_zstd_decode(cudaStream_t st, size_t n, const int32_t *methods, const uint8_t *src,
                   const uint64_t *src_off, const uint32_t *src_size, uint8_t *dst,
                   uint64_t dst_stride, uint32_t cap, uint32_t *out_size, int32_t *status,
                   uint8_t *scratch, const uint32_t *predef)
{
    if (zstd_use_cta_kernel())
        k_zstd_decode<<<(unsigned) n, ZSTDD_THREADS, ZSTDD_SMEM, st>>>(
            methods, src, src_off, src_size, dst, dst_stride, cap, out_size, status, scratch,
            ZSTDD_SCRATCH_BYTES);
    else
        k_zstd_decode_w<<<(unsigned) ((n + ZSW_WARPS - 1) / ZSW_WARPS), ZSW_THREADS, ZSW_SMEM, st>>>(
            methods, src, src_off, src_size, dst, dst_stride, cap, out_size, status, scratch,
            ZSTDD_SCRATCH_BYTES, predef, (uint32_t) n);
}

/* ----------------------------------------------------------------- context */

struct DevBuf
{
    void   *p = nullptr;
    size_t  cap = 0;
};

struct cryogpu_ctx
{
    int          device = -1;
    cudaStream_t stream = nullptr;      /* compute + copies of the *_host calls */
    cudaStream_t stream2 = nullptr;     /* second lane for double buffering */
    cudaEvent_t  ev[2] = {nullptr, nullptr};
    DevBuf       scratch;               /* per-block kernel scratch */
    uint32_t    *predef = nullptr;      /* predefined zstd FSE tables (device) */
    /* *_host staging (device + pinned host), two lanes */
    DevBuf       d_in[2], d_out[2], d_meta[2];
    DevBuf       h_in[2], h_meta[2], h_out[2];
    std::mutex   mu;
    bool         attrs_set = false;
    int          sm_count = 148;
};

static int
dev_reserve(DevBuf &b, size_t bytes)
{
    if (bytes <= b.cap)
        return CRYOGPU_OK;
    if (b.p)
        CU(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    bytes = (bytes + (1u << 20) - 1) & ~(size_t) ((1u << 20) - 1);
    CU(cudaMalloc(&b.p, bytes));
    b.cap = bytes;
    return CRYOGPU_OK;
}

static int
host_reserve(DevBuf &b, size_t bytes)
{
    if (bytes <= b.cap)
        return CRYOGPU_OK;
    if (b.p)
        CU(cudaFreeHost(b.p));
    b.p = nullptr;
    b.cap = 0;
    bytes = (bytes + (1u << 20) - 1) & ~(size_t) ((1u << 20) - 1);
    CU(cudaMallocHost(&b.p, bytes));
    b.cap = bytes;
    return CRYOGPU_OK;
}

static int
set_kernel_attrs(cryogpu_ctx *ctx)
{
    if (ctx->attrs_set)
        return CRYOGPU_OK;
    CU(cudaFuncSetAttribute(k_lz4_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, LZ4D_SMEM));
    CU(cudaFuncSetAttribute(k_lz4_decode_w, cudaFuncAttributeMaxDynamicSharedMemorySize, LZ4W_SMEM));
    CU(cudaFuncSetAttribute(k_zstd_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, ZSTDD_SMEM));
    CU(cudaFuncSetAttribute(k_zstd_decode_w, cudaFuncAttributeMaxDynamicSharedMemorySize, ZSW_SMEM));
    CU(cudaMalloc(&ctx->predef, ZSW_PREDEF_CELLS * sizeof(uint32_t)));
    k_zstd_build_predef<<<1, 32, 0, ctx->stream>>>(ctx->predef);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaFuncSetAttribute(k_lz4_encode, cudaFuncAttributeMaxDynamicSharedMemorySize, LZ4E_SMEM));
    CU(cudaFuncSetAttribute(k_zstd_encode, cudaFuncAttributeMaxDynamicSharedMemorySize, ZSTDE_SMEM));
    ctx->attrs_set = true;
    return CRYOGPU_OK;
}

extern "C" int
cryogpu_version(void)
{
    return CRYOGPU_VERSION;
}

extern "C" const char *
cryogpu_last_error(void)
{
    return g_err;
}

extern "C" const char *
cryogpu_status_string(int s)
{
    switch (s)
    {
        case CRYOGPU_ST_OK: return "ok";
        case CRYOGPU_ST_INPUT: return "truncated or over-long input";
        case CRYOGPU_ST_OUTPUT: return "output exceeds block capacity";
        case CRYOGPU_ST_OFFSET: return "match offset outside the output";
        case CRYOGPU_ST_FORMAT: return "malformed stream";
        case CRYOGPU_ST_SIZE: return "frame content size mismatch";
        case CRYOGPU_ST_METHOD: return "unknown compression method";
        case CRYOGPU_ST_UNSUPPORTED: return "unsupported stream feature";
        default: return "unknown status";
    }
}

extern "C" int
cryogpu_device_count(void)
{
    int n = 0;

    if (cudaGetDeviceCount(&n) != cudaSuccess)
    {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int
cryogpu_init(int device, cryogpu_ctx **out)
{
    if (!out)
        return fail(CRYOGPU_E_ARG, "cryogpu_init: ctx is NULL");
    *out = nullptr;
    int n = cryogpu_device_count();

    if (n <= 0)
        return fail(CRYOGPU_E_CUDA, "no CUDA device: libcryogpu has no CPU fallback");
    if (device < 0 || device >= n)
        return fail(CRYOGPU_E_ARG, "device %d out of range (0..%d)", device, n - 1);
    cudaDeviceProp prop;

    CU(cudaSetDevice(device));
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(CRYOGPU_E_CUDA, "device %d is sm_%d%d; libcryogpu is built for sm_100a only",
                    device, prop.major, prop.minor);
    cryogpu_ctx *ctx = new cryogpu_ctx();

    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 148;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev[1], cudaEventDisableTiming) != cudaSuccess)
    {
        delete ctx;
        return fail(CRYOGPU_E_CUDA, "stream/event creation failed: %s",
                    cudaGetErrorString(cudaGetLastError()));
    }
    int rc = set_kernel_attrs(ctx);

    if (rc != CRYOGPU_OK)
    {
        delete ctx;
        return rc;
    }
    *out = ctx;
    return CRYOGPU_OK;
}

extern "C" void
cryogpu_shutdown(cryogpu_ctx *ctx)
{
    if (!ctx)
        return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->stream2);
    cudaFree(ctx->scratch.p);
    cudaFree(ctx->predef);
    for (int i = 0; i < 2; i++)
    {
        cudaFree(ctx->d_in[i].p);
        cudaFree(ctx->d_out[i].p);
        cudaFree(ctx->d_meta[i].p);
        cudaFreeHost(ctx->h_in[i].p);
        cudaFreeHost(ctx->h_meta[i].p);
        cudaFreeHost(ctx->h_out[i].p);
        cudaEventDestroy(ctx->ev[i]);
    }
    cudaStreamDestroy(ctx->stream);
    cudaStreamDestroy(ctx->stream2);
    delete ctx;
}

extern "C" int
cryogpu_device(const cryogpu_ctx *ctx)
{
    return ctx ? ctx->device : -1;
}

extern "C" uint64_t
cryogpu_compress_bound(int method, uint64_t n)
{
    if (method == CRYOGPU_LZ4)
        return n + n / 255 + 16;                            /* LZ4_compressBound */
    if (method == CRYOGPU_ZSTD)                             /* ZSTD_compressBound */
        return n + (n >> 8) + (n < (128u << 10) ? (((128u << 10) - n) >> 11) : 0);
    return 0;
}

extern "C" void *
cryogpu_host_alloc(size_t bytes)
{
    void *p = nullptr;

    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess)
    {
        fail(CRYOGPU_E_NOMEM, "cudaMallocHost(%zu): %s", bytes,
             cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return p;
}

extern "C" void
cryogpu_host_free(void *p)
{
    if (p)
        cudaFreeHost(p);
}

/* ------------------------------------------------------- device-resident API */

extern "C" int
cryogpu_decompress_device(cryogpu_ctx *ctx, size_t n, const int32_t *d_methods,
                          const uint8_t *d_src, const uint64_t *d_src_off,
                          const uint32_t *d_src_size, uint8_t *d_dst, uint64_t dst_stride,
                          uint32_t block_size, uint32_t *d_out_size, int32_t *d_status,
                          void *stream)
{
    if (!ctx)
        return fail(CRYOGPU_E_ARG, "ctx is NULL");
    if (n == 0)
        return CRYOGPU_OK;
    if (!d_methods || !d_src || !d_src_off || !d_src_size || !d_dst || !d_out_size || !d_status)
        return fail(CRYOGPU_E_ARG, "NULL device pointer");
    if (((uintptr_t) d_dst & 15) || (dst_stride & 15) || dst_stride < block_size)
        return fail(CRYOGPU_E_ARG, "d_dst and dst_stride must be multiples of 16, stride >= block_size");
    if (block_size == 0 || block_size > (1u << 27) || n > 0x7fffffffu)
        return fail(CRYOGPU_E_ARG, "block_size or n out of range");
    cudaStream_t st = stream ? (cudaStream_t) stream : ctx->stream;

    CU(cudaSetDevice(ctx->device));
    {
        std::lock_guard<std::mutex> g(ctx->mu);
        int rc = dev_reserve(ctx->scratch, n * (size_t) ZSTDD_SCRATCH_BYTES);

        if (rc != CRYOGPU_OK)
            return rc;
    }
    k_flag_unknown_methods<<<(unsigned) ((n + 255) / 256), 256, 0, st>>>(d_methods, n, d_out_size,
                                                                         d_status);
    launch_lz4_decode(st, n, d_methods, d_src, d_src_off, d_src_size, d_dst, dst_stride, block_size,
                      d_out_size, d_status);
    launch_zstd_decode(st, n, d_methods, d_src, d_src_off, d_src_size, d_dst, dst_stride, block_size,
                       d_out_size, d_status, (uint8_t *) ctx->scratch.p, ctx->predef);
    CU(cudaGetLastError());
    return CRYOGPU_OK;
}

extern "C" int
cryogpu_compress_device(cryogpu_ctx *ctx, size_t n, int method, int level_or_accel,
                        const uint8_t *d_src, uint64_t src_stride, uint32_t block_size,
                        uint8_t *d_dst, uint64_t dst_stride, uint32_t dst_cap,
                        uint32_t *d_dst_size, int32_t *d_status, void *stream)
{
    if (!ctx)
        return fail(CRYOGPU_E_ARG, "ctx is NULL");
    if (method != CRYOGPU_LZ4 && method != CRYOGPU_ZSTD)
        return fail(CRYOGPU_E_METHOD, "unknown compression method %d", method);
    if (n == 0)
        return CRYOGPU_OK;
    if (!d_src || !d_dst || !d_dst_size || !d_status)
        return fail(CRYOGPU_E_ARG, "NULL device pointer");
    if (((uintptr_t) d_src & 15) || (src_stride & 15) || ((uintptr_t) d_dst & 15) || (dst_stride & 15))
        return fail(CRYOGPU_E_ARG, "device pointers and strides must be multiples of 16");
    if (block_size == 0 || block_size > (1u << 27) || src_stride < block_size || dst_stride < dst_cap ||
        n > 0x7fffffffu)
        return fail(CRYOGPU_E_ARG, "size out of range");
    if (dst_cap < cryogpu_compress_bound(method, block_size))
        return fail(CRYOGPU_E_ARG, "dst_cap %u below cryogpu_compress_bound", dst_cap);
    cudaStream_t st = stream ? (cudaStream_t) stream : ctx->stream;

    CU(cudaSetDevice(ctx->device));
    size_t per = method == CRYOGPU_LZ4 ? lz4e_scratch_bytes(block_size) : zstde_scratch_bytes(block_size);
    const size_t zgrid = std::min<size_t>(n, (size_t) ctx->sm_count);
    {
        std::lock_guard<std::mutex> g(ctx->mu);
        int rc = dev_reserve(ctx->scratch, (method == CRYOGPU_LZ4 ? n : zgrid) * per);

        if (rc != CRYOGPU_OK)
            return rc;
    }
    if (method == CRYOGPU_LZ4)
        k_lz4_encode<<<(unsigned) n, LZ4E_THREADS, LZ4E_SMEM, st>>>(
            d_src, src_stride, block_size, d_dst, dst_stride, dst_cap, level_or_accel, d_dst_size,
            d_status, (uint8_t *) ctx->scratch.p, per);
    else
        k_zstd_encode<<<(unsigned) zgrid, ZSTDE_THREADS, ZSTDE_SMEM, st>>>(
            d_src, src_stride, block_size, d_dst, dst_stride, dst_cap, level_or_accel, d_dst_size,
            d_status, (uint8_t *) ctx->scratch.p, per, (uint32_t) n);
    CU(cudaGetLastError());
    return CRYOGPU_OK;
}

/* -------------------------------------------------------------- host API */

/* blocks per pipeline chunk: bounds device/pinned staging to ~64 MiB per lane */
static size_t
chunk_blocks(uint32_t block_size)
{
    size_t c = (64u << 20) / block_size;

    return c < 1 ? 1 : c;
}

static bool
is_pinned(const void *p)
{
    cudaPointerAttributes a;

    if (cudaPointerGetAttributes(&a, p) != cudaSuccess)
    {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

extern "C" int
cryogpu_decompress_host(cryogpu_ctx *ctx, size_t n, const int32_t *methods,
                        const void *const *src, const uint32_t *src_size, void *const *dst,
                        uint32_t block_size, uint32_t *out_size, int32_t *status)
{
    if (!ctx)
        return fail(CRYOGPU_E_ARG, "ctx is NULL");
    if (n == 0)
        return CRYOGPU_OK;
    if (!methods || !src || !src_size || !dst || !out_size || !status)
        return fail(CRYOGPU_E_ARG, "NULL argument");
    if (block_size == 0 || block_size > (1u << 27))
        return fail(CRYOGPU_E_ARG, "block_size out of range");
    std::lock_guard<std::mutex> g(ctx->mu);

    CU(cudaSetDevice(ctx->device));
    const size_t   chunk = chunk_blocks(block_size);
    const uint64_t stride = ((uint64_t) block_size + 15) & ~(uint64_t) 15;
    const bool     dst_pinned = is_pinned(dst[0]);
    cudaStream_t   lanes[2] = {ctx->stream, ctx->stream2};
    size_t         pending_lo[2] = {0, 0}, pending_n[2] = {0, 0};

    /* finish a lane: wait, then hand results to the caller */
    auto drain = [&](int l) -> int {
        if (pending_n[l] == 0)
            return CRYOGPU_OK;
        CU(cudaStreamSynchronize(lanes[l]));
        size_t   lo = pending_lo[l], cnt = pending_n[l];
        uint8_t *hm = (uint8_t *) ctx->h_meta[l].p;
        uint32_t *h_osz = (uint32_t *) (hm + cnt * 16);
        int32_t  *h_st = (int32_t *) (hm + cnt * 20);

        memcpy(out_size + lo, h_osz, cnt * 4);
        memcpy(status + lo, h_st, cnt * 4);
        if (!dst_pinned)
            for (size_t i = 0; i < cnt; i++)
                memcpy(dst[lo + i], (uint8_t *) ctx->h_out[l].p + i * stride, block_size);
        pending_n[l] = 0;
        return CRYOGPU_OK;
    };

    int lane = 0;

    for (size_t lo = 0; lo < n; lo += chunk, lane ^= 1)
    {
        size_t cnt = std::min(chunk, n - lo);
        int    rc = drain(lane);

        if (rc != CRYOGPU_OK)
            return rc;
        /* meta layout (host and device): off u64[cnt] | size u32[cnt] | method i32[cnt] ->
         * out_size u32[cnt] | status i32[cnt]   (offsets cnt*{0,8,12,16,20}) */
        size_t in_bytes = 0;

        for (size_t i = 0; i < cnt; i++)
            in_bytes += ((size_t) src_size[lo + i] + 15) & ~(size_t) 15;
        if ((rc = host_reserve(ctx->h_in[lane], in_bytes + 16)) != CRYOGPU_OK ||
            (rc = host_reserve(ctx->h_meta[lane], cnt * 24)) != CRYOGPU_OK ||
            (rc = dev_reserve(ctx->d_in[lane], in_bytes + 16)) != CRYOGPU_OK ||
            (rc = dev_reserve(ctx->d_meta[lane], cnt * 24)) != CRYOGPU_OK ||
            (rc = dev_reserve(ctx->d_out[lane], cnt * stride)) != CRYOGPU_OK ||
            (rc = dev_reserve(ctx->scratch, chunk * (size_t) ZSTDD_SCRATCH_BYTES * 2)) != CRYOGPU_OK)
            return rc;
        if (!dst_pinned && (rc = host_reserve(ctx->h_out[lane], cnt * stride)) != CRYOGPU_OK)
            return rc;
        uint8_t  *hm = (uint8_t *) ctx->h_meta[lane].p;
        uint64_t *h_off = (uint64_t *) hm;
        uint32_t *h_sz = (uint32_t *) (hm + cnt * 8);
        int32_t  *h_me = (int32_t *) (hm + cnt * 12);
        size_t    p = 0;

        for (size_t i = 0; i < cnt; i++)
        {
            h_off[i] = p;
            h_sz[i] = src_size[lo + i];
            h_me[i] = methods[lo + i];
            if (src_size[lo + i])
                memcpy((uint8_t *) ctx->h_in[lane].p + p, src[lo + i], src_size[lo + i]);
            p += ((size_t) src_size[lo + i] + 15) & ~(size_t) 15;
        }
        cudaStream_t st = lanes[lane];
        uint8_t     *dm = (uint8_t *) ctx->d_meta[lane].p;

        CU(cudaMemcpyAsync(ctx->d_in[lane].p, ctx->h_in[lane].p, in_bytes + 16, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(dm, hm, cnt * 16, cudaMemcpyHostToDevice, st));
        uint8_t *scr = (uint8_t *) ctx->scratch.p + (size_t) lane * chunk * ZSTDD_SCRATCH_BYTES;

        k_flag_unknown_methods<<<(unsigned) ((cnt + 255) / 256), 256, 0, st>>>(
            (int32_t *) (dm + cnt * 12), cnt, (uint32_t *) (dm + cnt * 16), (int32_t *) (dm + cnt * 20));
        launch_lz4_decode(st, cnt, (int32_t *) (dm + cnt * 12), (uint8_t *) ctx->d_in[lane].p,
                          (uint64_t *) dm, (uint32_t *) (dm + cnt * 8),
                          (uint8_t *) ctx->d_out[lane].p, stride, block_size,
                          (uint32_t *) (dm + cnt * 16), (int32_t *) (dm + cnt * 20));
        launch_zstd_decode(st, cnt, (int32_t *) (dm + cnt * 12), (uint8_t *) ctx->d_in[lane].p,
                           (uint64_t *) dm, (uint32_t *) (dm + cnt * 8),
                           (uint8_t *) ctx->d_out[lane].p, stride, block_size,
                           (uint32_t *) (dm + cnt * 16), (int32_t *) (dm + cnt * 20), scr,
                           ctx->predef);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(hm + cnt * 16, dm + cnt * 16, cnt * 8, cudaMemcpyDeviceToHost, st));
        if (dst_pinned)
        {
            /* straight into the caller's pinned blocks; merge contiguous runs */
            size_t i = 0;

            while (i < cnt)
            {
                size_t j = i + 1;

                while (j < cnt && stride == block_size &&
                       (uint8_t *) dst[lo + j] == (uint8_t *) dst[lo + j - 1] + block_size)
                    j++;
                CU(cudaMemcpyAsync(dst[lo + i], (uint8_t *) ctx->d_out[lane].p + i * stride,
                                   (j - i - 1) * stride + block_size, cudaMemcpyDeviceToHost, st));
                i = j;
            }
        }
        else
            CU(cudaMemcpyAsync(ctx->h_out[lane].p, ctx->d_out[lane].p, cnt * stride,
                               cudaMemcpyDeviceToHost, st));
        pending_lo[lane] = lo;
        pending_n[lane] = cnt;
    }
    int rc = drain(0);

    if (rc == CRYOGPU_OK)
        rc = drain(1);
    return rc;
}

extern "C" int
cryogpu_compress_host(cryogpu_ctx *ctx, size_t n, int method, int level_or_accel,
                      const void *const *src, uint32_t block_size, void *const *dst,
                      uint32_t dst_cap, uint32_t *dst_size, int32_t *status)
{
    if (!ctx)
        return fail(CRYOGPU_E_ARG, "ctx is NULL");
    if (method != CRYOGPU_LZ4 && method != CRYOGPU_ZSTD)
        return fail(CRYOGPU_E_METHOD, "unknown compression method %d", method);
    if (n == 0)
        return CRYOGPU_OK;
    if (!src || !dst || !dst_size || !status)
        return fail(CRYOGPU_E_ARG, "NULL argument");
    if (block_size == 0 || block_size > (1u << 27))
        return fail(CRYOGPU_E_ARG, "block_size out of range");
    const uint64_t bound = cryogpu_compress_bound(method, block_size);

    if (dst_cap < bound)
        return fail(CRYOGPU_E_ARG, "dst_cap %u below cryogpu_compress_bound %llu", dst_cap,
                    (unsigned long long) bound);
    std::lock_guard<std::mutex> g(ctx->mu);

    CU(cudaSetDevice(ctx->device));
    const size_t   chunk = chunk_blocks(block_size);
    const uint64_t sstride = ((uint64_t) block_size + 15) & ~(uint64_t) 15;
    const uint64_t dstride = (bound + 15) & ~(uint64_t) 15;
    const size_t   per = method == CRYOGPU_LZ4 ? lz4e_scratch_bytes(block_size)
                                               : zstde_scratch_bytes(block_size);
    const bool     src_pinned = is_pinned(src[0]);
    cudaStream_t   lanes[2] = {ctx->stream, ctx->stream2};
    size_t         pending_lo[2] = {0, 0}, pending_n[2] = {0, 0};

    auto drain = [&](int l) -> int {
        if (pending_n[l] == 0)
            return CRYOGPU_OK;
        CU(cudaStreamSynchronize(lanes[l]));
        size_t    lo = pending_lo[l], cnt = pending_n[l];
        uint32_t *h_sz = (uint32_t *) ctx->h_meta[l].p;
        int32_t  *h_st = (int32_t *) ((uint8_t *) ctx->h_meta[l].p + cnt * 4);

        for (size_t i = 0; i < cnt; i++)
        {
            dst_size[lo + i] = h_sz[i];
            status[lo + i] = h_st[i];
            if (h_st[i] == CRYOGPU_ST_OK)
                memcpy(dst[lo + i], (uint8_t *) ctx->h_out[l].p + i * dstride, h_sz[i]);
        }
        pending_n[l] = 0;
        return CRYOGPU_OK;
    };

    int lane = 0;

    for (size_t lo = 0; lo < n; lo += chunk, lane ^= 1)
    {
        size_t cnt = std::min(chunk, n - lo);
        int    rc = drain(lane);

        if (rc != CRYOGPU_OK)
            return rc;
        if ((rc = dev_reserve(ctx->d_in[lane], cnt * sstride)) != CRYOGPU_OK ||
            (rc = dev_reserve(ctx->d_out[lane], cnt * dstride)) != CRYOGPU_OK ||
            (rc = dev_reserve(ctx->d_meta[lane], cnt * 8)) != CRYOGPU_OK ||
            (rc = host_reserve(ctx->h_meta[lane], cnt * 8)) != CRYOGPU_OK ||
            (rc = host_reserve(ctx->h_out[lane], cnt * dstride)) != CRYOGPU_OK ||
            (rc = dev_reserve(ctx->scratch, 2 * chunk * per)) != CRYOGPU_OK)
            return rc;
        if (!src_pinned && (rc = host_reserve(ctx->h_in[lane], cnt * sstride)) != CRYOGPU_OK)
            return rc;
        cudaStream_t st = lanes[lane];

        if (src_pinned)
            for (size_t i = 0; i < cnt; i++)
                CU(cudaMemcpyAsync((uint8_t *) ctx->d_in[lane].p + i * sstride, src[lo + i], block_size,
                                   cudaMemcpyHostToDevice, st));
        else
        {
            for (size_t i = 0; i < cnt; i++)
                memcpy((uint8_t *) ctx->h_in[lane].p + i * sstride, src[lo + i], block_size);
            CU(cudaMemcpyAsync(ctx->d_in[lane].p, ctx->h_in[lane].p, cnt * sstride,
                               cudaMemcpyHostToDevice, st));
        }
        uint8_t *dm = (uint8_t *) ctx->d_meta[lane].p;
        uint8_t *scr = (uint8_t *) ctx->scratch.p + (size_t) lane * chunk * per;

        if (method == CRYOGPU_LZ4)
            k_lz4_encode<<<(unsigned) cnt, LZ4E_THREADS, LZ4E_SMEM, st>>>(
                (uint8_t *) ctx->d_in[lane].p, sstride, block_size, (uint8_t *) ctx->d_out[lane].p,
                dstride, (uint32_t) bound, level_or_accel, (uint32_t *) dm, (int32_t *) (dm + cnt * 4),
                scr, per);
        else
            k_zstd_encode<<<(unsigned) std::min<size_t>(cnt, (size_t) ctx->sm_count), ZSTDE_THREADS,
                            ZSTDE_SMEM, st>>>(
                (uint8_t *) ctx->d_in[lane].p, sstride, block_size, (uint8_t *) ctx->d_out[lane].p,
                dstride, (uint32_t) bound, level_or_accel, (uint32_t *) dm, (int32_t *) (dm + cnt * 4),
                scr, per, (uint32_t) cnt);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(ctx->h_meta[lane].p, dm, cnt * 8, cudaMemcpyDeviceToHost, st));
        /* compressed sizes are unknown until the kernel ends: bring back whole slots when the
         * batch is tiny, otherwise sizes first and then only the used prefix of every slot */
        if (cnt <= 4)
            CU(cudaMemcpyAsync(ctx->h_out[lane].p, ctx->d_out[lane].p, cnt * dstride,
                               cudaMemcpyDeviceToHost, st));
        else
        {
            CU(cudaStreamSynchronize(st));
            uint32_t *h_sz = (uint32_t *) ctx->h_meta[lane].p;

            for (size_t i = 0; i < cnt; i++)
                CU(cudaMemcpyAsync((uint8_t *) ctx->h_out[lane].p + i * dstride,
                                   (uint8_t *) ctx->d_out[lane].p + i * dstride,
                                   std::min<uint64_t>(h_sz[i], dstride), cudaMemcpyDeviceToHost, st));
        }
        pending_lo[lane] = lo;
        pending_n[lane] = cnt;
    }
    int rc = drain(0);

    if (rc == CRYOGPU_OK)
        rc = drain(1);
    return rc;
}

/* ------------------------------------------------------------- multi GPU */

template <typename F>
static int
run_sharded(int nctx, size_t n, F fn)
{
    std::vector<std::thread> th;
    std::vector<int>         rc(nctx, CRYOGPU_OK);
    std::vector<std::string> msg(nctx);

    for (int k = 0; k < nctx; k++)
    {
        size_t lo = n * k / nctx, hi = n * (k + 1) / nctx;

        if (hi == lo)
            continue;
        th.emplace_back([&, k, lo, hi]() {
            rc[k] = fn(k, lo, hi - lo);
            if (rc[k] != CRYOGPU_OK)
                msg[k] = g_err;
        });
    }
    for (auto &t : th)
        t.join();
    for (int k = 0; k < nctx; k++)
        if (rc[k] != CRYOGPU_OK)
            return fail(rc[k], "gpu shard %d: %s", k, msg[k].c_str());
    return CRYOGPU_OK;
}

extern "C" int
cryogpu_decompress_host_multi(cryogpu_ctx *const *ctxs, int nctx, size_t n, const int32_t *methods,
                              const void *const *src, const uint32_t *src_size, void *const *dst,
                              uint32_t block_size, uint32_t *out_size, int32_t *status)
{
    if (!ctxs || nctx < 1)
        return fail(CRYOGPU_E_ARG, "no contexts");
    return run_sharded(nctx, n, [&](int k, size_t lo, size_t cnt) {
        return cryogpu_decompress_host(ctxs[k], cnt, methods + lo, src + lo, src_size + lo, dst + lo,
                                       block_size, out_size + lo, status + lo);
    });
}

extern "C" int
cryogpu_compress_host_multi(cryogpu_ctx *const *ctxs, int nctx, size_t n, int method,
                            int level_or_accel, const void *const *src, uint32_t block_size,
                            void *const *dst, uint32_t dst_cap, uint32_t *dst_size, int32_t *status)
{
    if (!ctxs || nctx < 1)
        return fail(CRYOGPU_E_ARG, "no contexts");
    return run_sharded(nctx, n, [&](int k, size_t lo, size_t cnt) {
        return cryogpu_compress_host(ctxs[k], cnt, method, level_or_accel, src + lo, block_size,
                                     dst + lo, dst_cap, dst_size + lo, status + lo);
    });
}
