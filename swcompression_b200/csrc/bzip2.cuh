// bzip2.cuh — argument block of the BZip2 kernel.
#pragma once
#include "common.cuh"

namespace swc {
namespace bzip2 {

struct Args {
    const u8 *in_base;
    const u64 *in_off, *in_len;
    u8 *out_base;
    const u64 *out_off, *out_cap;
    u64 *out_len, *consumed_bits;
    int32_t *status;
    u64 n;
    u8 *scratch;
    const u64 *scr_off;      // per unit, bytes into scratch
};

size_t scratch_per_unit(u64 out_cap);
int launch(const Args &a, cudaStream_t stream);

}  // namespace bzip2
}  // namespace swc
