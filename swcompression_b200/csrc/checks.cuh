// checks.cuh — device checksum entry points (all pointers are device pointers; results are 64-bit slots).
#pragma once
#include <vector>
#include "common.cuh"

namespace swc {
namespace checks {

size_t partial_bytes(u64 n);   // scratch for the per-chunk partials
int crc32(const u8 *d, u64 n, u64 *d_result, u64 *d_partial, cudaStream_t s);
int bzip2_crc32(const u8 *d, u64 n, u64 *d_result, u64 *d_partial, cudaStream_t s);
int crc64(const u8 *d, u64 n, u64 *d_result, u64 *d_partial, cudaStream_t s);
int adler32(const u8 *d, u64 n, u64 *d_result, u64 *d_partial, cudaStream_t s);
int xxh32_batch(const u8 *base, const u64 *off /* may be null */, const u64 *len, u32 *result, u64 n, cudaStream_t s);
int sha256(const u8 *d, u64 n, u8 *d_digest, cudaStream_t s);
int crc32_units(const u8 *base, const u64 *off, const u64 *len, const int32_t *status, u32 *result, u64 n, cudaStream_t s);   // warp per buffer; units with status != 0 are skipped (status may be null)
int find_gzip_members(const u8 *d_in, u64 n, std::vector<size_t> &pos);   // sorted candidate member starts (device scan)
int find_bzip2_magics(const u8 *d_in, u64 begin, u64 n, std::vector<u64> &entries);   // sorted (bit position << 1 | is_end_magic), device scan
int gather_units(const u8 *src, const u64 *src_off, const u64 *len, u8 *dst, const u64 *dst_off, u64 n, cudaStream_t s);

// host convenience: run a 32/64-bit check over a device buffer and fetch the value (blocking)
enum Kind { CRC32 = 0, BZIP2_CRC32 = 1, CRC64 = 2, ADLER32 = 3, XXH32 = 4 };
int check_device(Kind k, const u8 *d, u64 n, u64 *value);

}  // namespace checks
}  // namespace swc
