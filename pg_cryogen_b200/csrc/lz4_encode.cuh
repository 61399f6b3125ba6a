/* lz4_encode.cuh -- placeholder until the LZ4 block encoder lands */
#pragma once
#include "cryo_common.cuh"
#define LZ4E_THREADS 128
#define LZ4E_SMEM 1024
static inline size_t lz4e_scratch_bytes(uint32_t block_size) { return 1024; }
CRYO_DEV void lz4_encode_block(const uint8_t *src, uint32_t n, uint8_t *dst, uint32_t dst_cap, int accel,
                               uint32_t *dst_size, int32_t *status, uint8_t *scratch)
{
    if (threadIdx.x == 0) { *dst_size = 0; *status = ST_UNSUPPORTED; }
}
