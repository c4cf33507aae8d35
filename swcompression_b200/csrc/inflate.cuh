// inflate.cuh — argument block shared by the Deflate kernels and the host launcher.
#pragma once
#include "common.cuh"

namespace swc {
namespace inflate {

struct BatchArgs {
    const u8 *in_base;
    const u64 *in_off, *in_len;
    const u8 *start_bits;      // may be null
    u8 *out_base;
    const u64 *out_off, *out_cap;
    u64 *out_len, *consumed_bits;
    int32_t *status;
    u64 n;
    u32 *rec_base;             // match-record scratch
    u32 *rec_count;            // n entries
    unsigned long long *ticket; // unit ticket counters of the persistent-lane kernels: [0] K1 / K1w, [1] K1L (zeroed by the launcher)
};

// A unit whose output region starts at byte `out_off` owns records [out_off/3, (out_off+cap)/3): every record accounts
// for >= 3 output bytes, so disjoint output regions give disjoint record regions without a prefix sum.
__host__ __device__ __forceinline__ u64 rec_start(u64 out_off) { return out_off / 3; }
inline size_t scratch_bytes(u64 n, u64 out_capacity_total) {
    return (size_t)((out_capacity_total / 3 + 2) * 4 + n * 4 + 1024);
}

int launch(const BatchArgs &a, cudaStream_t stream);
void launch_slow(const BatchArgs &a, cudaStream_t stream);
int launch_warp(const BatchArgs &a, cudaStream_t stream);    // inflate_warp.cu (K1w)
int launch_lut(const BatchArgs &a, cudaStream_t stream);     // inflate_lut.cu  (K1L: table-lookup decode, thread per unit)

}  // namespace inflate
}  // namespace swc
