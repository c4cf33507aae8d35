// common.cuh — shared device/host helpers for libswcgpu (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/swc_status.h"

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int64_t i64;

#define SWC_WARP 32
#define SWC_FULL 0xFFFFFFFFu

// internal status: unit needs the generic (slow) decoder — never escapes the library
#define SWC_INTERNAL_NEEDS_SLOW 9001

namespace swc {

__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31; }

// launch bookkeeping (swc_kernel_launches)
void count_launch(int n = 1);
// record a CUDA failure; returns SWC_ERR_CUDA
int cuda_fail(cudaError_t e, const char *where);

#define SWC_CUDA_TRY(expr)                                                     \
    do {                                                                       \
        cudaError_t _e = (expr);                                               \
        if (_e != cudaSuccess) return swc::cuda_fail(_e, #expr);               \
    } while (0)

// optional per-kernel timing: when enabled, launchers drop a CUDA event on `stream` between their kernels
void timing_mark(cudaStream_t stream);

// scratch pool: grow-only per-device buffer, used when the caller passes no scratch
int scratch_get(size_t bytes, void **p, cudaStream_t stream);
// grow-only arenas reused across calls: slot 0 = kernel scratch, 1..3 = device staging of the *_batch_host paths
int arena_get(int slot, size_t bytes, void **p, cudaStream_t stream);

}  // namespace swc
