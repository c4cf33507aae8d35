// checks.cu — device-side checksums used by the wrappers (reference: Sources/Common/CheckSums.swift:12-57,
// Sources/LZ4/XxHash32.swift:24-83, Sources/XZ/Sha256.swift).
//
// CRC-32 / CRC-32(bzip2) / CRC-64 are linear over GF(2): every thread folds one 4 KiB chunk with a byte table held in
// shared memory, then one thread chains the chunk values with x^(8*len) multiplications (zlib's crc32_combine
// identity, which holds for the conditioned CRC values).  Adler-32 splits the same way.  xxHash32 and SHA-256 have no
// combine operator, so they run one thread per buffer (many buffers in parallel for LZ4 block checksums).
#include "common.cuh"
#include <algorithm>
#include <vector>
#include "checks.cuh"

namespace swc {
namespace checks {

constexpr u32 CHUNK = 4096;

// ---------------------------------------------------------------- GF(2) helpers (device)
// reflected 32-bit: bit 31 = x^0
__device__ u32 mulmod32r(u32 a, u32 b) {
    u32 m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) { p ^= b; if ((a & (m - 1)) == 0) break; }
        m >>= 1;
        b = (b & 1) ? (b >> 1) ^ 0xEDB88320u : b >> 1;
    }
    return p;
}
__device__ u64 mulmod64r(u64 a, u64 b) {
    u64 m = 1ull << 63, p = 0;
    for (;;) {
        if (a & m) { p ^= b; if ((a & (m - 1)) == 0) break; }
        m >>= 1;
        b = (b & 1) ? (b >> 1) ^ 0xC96C5795D7870F42ull : b >> 1;
    }
    return p;
}
// normal (MSB-first) 32-bit: bit 0 = x^0, poly 0x04C11DB7
__device__ u32 mulmod32n(u32 a, u32 b) {
    u32 p = 0;
    while (a) {
        if (a & 1) p ^= b;
        a >>= 1;
        b = (b & 0x80000000u) ? (b << 1) ^ 0x04C11DB7u : b << 1;
    }
    return p;
}
// x^(8*nbytes) mod P in each representation
__device__ u32 xpow32r(u64 nbytes) {
    u32 p = 1u << 31, sq = 1u << 23;           // x^0, x^8
    while (nbytes) { if (nbytes & 1) p = mulmod32r(sq, p); sq = mulmod32r(sq, sq); nbytes >>= 1; }
    return p;
}
__device__ u64 xpow64r(u64 nbytes) {
    u64 p = 1ull << 63, sq = 1ull << 55;
    while (nbytes) { if (nbytes & 1) p = mulmod64r(sq, p); sq = mulmod64r(sq, sq); nbytes >>= 1; }
    return p;
}
__device__ u32 xpow32n(u64 nbytes) {
    u32 p = 1, sq = 1u << 8;
    while (nbytes) { if (nbytes & 1) p = mulmod32n(sq, p); sq = mulmod32n(sq, sq); nbytes >>= 1; }
    return p;
}

// ---------------------------------------------------------------- chunk kernels
template <int KIND>   // 0 crc32, 1 bzip2 crc32, 2 crc64, 3 adler32
__global__ void __launch_bounds__(128) chunk_kernel(const u8 *data, u64 n, u64 *partial) {
    __shared__ u64 tab[256];
    for (u32 i = threadIdx.x; i < 256; i += blockDim.x) {
        if (KIND == 0) { u32 c = i; for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1; tab[i] = c; }
        if (KIND == 1) { u32 c = i << 24; for (int k = 0; k < 8; k++) c = (c & 0x80000000u) ? (c << 1) ^ 0x04C11DB7u : c << 1; tab[i] = c; }
        if (KIND == 2) { u64 c = i; for (int k = 0; k < 8; k++) c = (c & 1) ? 0xC96C5795D7870F42ull ^ (c >> 1) : c >> 1; tab[i] = c; }
    }
    __syncthreads();
    const u64 chunk = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 beg = chunk * CHUNK;
    if (beg >= n) return;
    const u64 end = beg + CHUNK < n ? beg + CHUNK : n;
    if (KIND == 0) {
        u32 c = 0xFFFFFFFFu;
        for (u64 i = beg; i < end; i++) c = (u32)tab[(c ^ data[i]) & 0xFF] ^ (c >> 8);
        partial[chunk] = ~c;
    } else if (KIND == 1) {
        u32 c = 0xFFFFFFFFu;
        for (u64 i = beg; i < end; i++) c = (c << 8) ^ (u32)tab[((c >> 24) ^ data[i]) & 0xFF];
        partial[chunk] = ~c;
    } else if (KIND == 2) {
        u64 c = ~0ull;
        for (u64 i = beg; i < end; i++) c = tab[(c ^ data[i]) & 0xFF] ^ (c >> 8);
        partial[chunk] = ~c;
    } else {
        u32 s1 = 0, s2 = 0;                       // sums of this chunk alone (s1 without the leading 1)
        for (u64 i = beg; i < end; i++) { s1 += data[i]; s2 += s1; }   // 4096*255 and 4096*4097/2*255 fit in u32
        partial[chunk] = (u64)(s1 % 65521u) | ((u64)(s2 % 65521u) << 32);
    }
}

template <int KIND>
__global__ void combine_kernel(const u64 *partial, u64 n, u64 *result) {
    const u64 nchunks = (n + CHUNK - 1) / CHUNK;
    if (KIND == 3) {
        u64 a = 1, b = 0;                         // CheckSums.swift:48-57
        for (u64 c = 0; c < nchunks; c++) {
            u64 len = (c + 1) * CHUNK <= n ? CHUNK : n - c * CHUNK;
            u64 s1 = partial[c] & 0xFFFFFFFFu, s2 = partial[c] >> 32;
            b = (b + (len % 65521u) * a + s2) % 65521u;
            a = (a + s1) % 65521u;
        }
        *result = (b << 16) + a;
        return;
    }
    if (nchunks == 0) { *result = 0; return; }    // CRC of the empty string is 0 in all three variants
    u64 acc = partial[0];
    const u32 p32r = KIND == 0 ? xpow32r(CHUNK) : 0;
    const u32 p32n = KIND == 1 ? xpow32n(CHUNK) : 0;
    const u64 p64r = KIND == 2 ? xpow64r(CHUNK) : 0;
    for (u64 c = 1; c < nchunks; c++) {
        const bool full = (c + 1) * CHUNK <= n;
        const u64 len = full ? CHUNK : n - c * CHUNK;
        if (KIND == 0) acc = mulmod32r(full ? p32r : xpow32r(len), (u32)acc) ^ (u32)partial[c];
        if (KIND == 1) acc = mulmod32n(full ? p32n : xpow32n(len), (u32)acc) ^ (u32)partial[c];
        if (KIND == 2) acc = mulmod64r(full ? p64r : xpow64r(len), acc) ^ partial[c];
    }
    *result = acc;
}

template <int KIND>
static int run(const u8 *d_data, u64 n, u64 *d_result, u64 *d_partial, cudaStream_t s) {
    const u64 nchunks = (n + CHUNK - 1) / CHUNK;
    if (nchunks) {
        chunk_kernel<KIND><<<(unsigned)((nchunks + 127) / 128), 128, 0, s>>>(d_data, n, d_partial);
        count_launch();
    }
    combine_kernel<KIND><<<1, 1, 0, s>>>(d_partial, n, d_result);
    count_launch();
    SWC_CUDA_TRY(cudaGetLastError());
    return SWC_OK;
}

size_t partial_bytes(u64 n) { return (size_t)(((n + CHUNK - 1) / CHUNK) + 1) * 8; }

int crc32(const u8 *d, u64 n, u64 *d_result, u64 *d_partial, cudaStream_t s) { return run<0>(d, n, d_result, d_partial, s); }
int bzip2_crc32(const u8 *d, u64 n, u64 *d_result, u64 *d_partial, cudaStream_t s) { return run<1>(d, n, d_result, d_partial, s); }
int crc64(const u8 *d, u64 n, u64 *d_result, u64 *d_partial, cudaStream_t s) { return run<2>(d, n, d_result, d_partial, s); }
int adler32(const u8 *d, u64 n, u64 *d_result, u64 *d_partial, cudaStream_t s) { return run<3>(d, n, d_result, d_partial, s); }

// ---------------------------------------------------------------- CRC-32 of many buffers: one warp per buffer
// Each lane folds one of 32 contiguous segments (a multiple of 16 bytes, read with 16-byte loads) four bytes per step with
// four byte tables in shared memory ("slicing by 4": 2.8 instead of 6 instructions per byte), then the 32 conditioned values
// are chained with x^(8*len) multiplications — the same identity as above, inside one warp.
__device__ __forceinline__ u32 crc_word(const u32 (*tab)[256], u32 c, u32 w) {
    c ^= w;
    return tab[3][c & 0xFF] ^ tab[2][(c >> 8) & 0xFF] ^ tab[1][(c >> 16) & 0xFF] ^ tab[0][c >> 24];
}
__global__ void __launch_bounds__(256) crc32_units_kernel(const u8 *base, const u64 *off, const u64 *len, const int32_t *status, u32 *result, u64 n) {
    __shared__ u32 tab[4][256];
    for (u32 i = threadIdx.x; i < 256; i += blockDim.x) {
        u32 c = i;
        for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        tab[0][i] = c;
    }
    __syncthreads();
    for (int t = 1; t < 4; t++) {
        for (u32 i = threadIdx.x; i < 256; i += blockDim.x) { const u32 c = tab[t - 1][i]; tab[t][i] = (c >> 8) ^ tab[0][c & 0xFF]; }
        __syncthreads();
    }
    const u64 unit = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (unit >= n) return;
    const u32 lane = threadIdx.x & 31;
    const u8 *p = base + off[unit];
    // a unit that did not decode cleanly has no defined extent (Deflate reports the size it WOULD need on overflow): skip it
    const u64 L = (status && status[unit] != SWC_OK) ? 0 : len[unit];
    const u64 seg = (((L + 31) / 32) + 15) & ~(u64)15;
    const u64 sb = lane * seg < L ? lane * seg : L, se = sb + seg < L ? sb + seg : L;
    u32 c = 0xFFFFFFFFu;
    u64 i = sb;
    for (; i < se && ((uintptr_t)(p + i) & 15); i++) c = tab[0][(c ^ p[i]) & 0xFF] ^ (c >> 8);
    for (; i + 16 <= se; i += 16) {
        const uint4 v = __ldg((const uint4 *)(p + i));
        c = crc_word(tab, c, v.x); c = crc_word(tab, c, v.y); c = crc_word(tab, c, v.z); c = crc_word(tab, c, v.w);
    }
    for (; i < se; i++) c = tab[0][(c ^ p[i]) & 0xFF] ^ (c >> 8);
    c = ~c;
    const u32 pw = xpow32r(seg);
    u32 acc = __shfl_sync(0xFFFFFFFFu, c, 0);
    for (int l = 1; l < 32; l++) {
        const u32 pl = __shfl_sync(0xFFFFFFFFu, c, l);
        const u64 lb = (u64)l * seg < L ? (u64)l * seg : L, le = lb + seg < L ? lb + seg : L;
        if (le == lb) continue;
        acc = mulmod32r(le - lb == seg ? pw : xpow32r(le - lb), acc) ^ pl;
    }
    if (lane == 0) result[unit] = acc;
}

int crc32_units(const u8 *base, const u64 *off, const u64 *len, const int32_t *status, u32 *result, u64 n, cudaStream_t s) {
    if (!n) return SWC_OK;
    crc32_units_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, s>>>(base, off, len, status, result, n);
    count_launch();
    SWC_CUDA_TRY(cudaGetLastError());
    return SWC_OK;
}

// gzip member signature scan (GzipHeader.swift:70-81: magic 1f 8b, method 8, reserved flag bits clear); positions are
// appended in any order and sorted by the host wrapper
__global__ void __launch_bounds__(256) gzip_scan_kernel(const u8 *d, u64 n, u64 *list, unsigned long long *count, u64 cap) {
    const u64 base = ((u64)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (base >= n) return;
    // 16 positions per thread from one aligned 16-byte load plus the 3 bytes that follow
    const uint4 v = *(const uint4 *)(d + base);                       // the buffer is padded by >= 32 bytes
    const u32 w[5] = {v.x, v.y, v.z, v.w, *(const u32 *)(d + base + 16)};
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const u32 x = __funnelshift_r(w[i >> 2], w[(i >> 2) + 1], (i & 3) * 8);
        if ((x & 0xE0FFFFFFu) == 0x00088B1Fu && base + i + 20 <= n) {
            const u64 slot = atomicAdd(count, 1ull);
            if (slot < cap) list[slot] = base + i;
        }
    }
}

int find_gzip_members(const u8 *d_in, u64 n, std::vector<size_t> &pos) {
    pos.clear();
    if (n < 20) return SWC_OK;
    const u64 cap = 1u << 22;
    void *p = nullptr;
    int st = arena_get(2, 256 + cap * 8, &p, 0);
    if (st) return st;
    unsigned long long *count = (unsigned long long *)p;
    u64 *list = (u64 *)((u8 *)p + 256);
    SWC_CUDA_TRY(cudaMemsetAsync(count, 0, 8, 0));
    const u64 threads = (n + 15) / 16;
    gzip_scan_kernel<<<(unsigned)((threads + 255) / 256), 256>>>(d_in, n, list, count, cap);
    count_launch();
    SWC_CUDA_TRY(cudaGetLastError());
    unsigned long long h = 0;
    SWC_CUDA_TRY(cudaMemcpy(&h, count, 8, cudaMemcpyDeviceToHost));
    if (h > cap) return SWC_OK;                                       // absurd candidate count: caller walks sequentially
    std::vector<u64> tmp(h);
    if (h) SWC_CUDA_TRY(cudaMemcpy(tmp.data(), list, h * 8, cudaMemcpyDeviceToHost));
    std::sort(tmp.begin(), tmp.end());
    pos.assign(tmp.begin(), tmp.end());
    return SWC_OK;
}

// bzip2 block / end-of-stream magics at EVERY bit offset (blocks are not byte aligned: BZip2.swift:71-90 reads 48 bits from
// wherever the previous block ended).  Entry = bit position << 1 | (1 for the end-of-stream magic).
__global__ void __launch_bounds__(256) bzip2_magic_scan_kernel(const u8 *d, u64 begin, u64 n, u64 *list, unsigned long long *count, u64 cap) {
    const u64 i = begin + (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 6 > n) return;                                            // a magic needs 48 bits
    u64 w = 0;
#pragma unroll
    for (int k = 0; k < 7; k++) w = (w << 8) | (i + k < n ? d[i + k] : 0);   // bits of bytes i..i+6, MSB first
#pragma unroll
    for (int b = 0; b < 8; b++) {
        if (i * 8 + b + 48 > n * 8) break;
        const u64 v = (w >> (8 - b)) & 0xFFFFFFFFFFFFull;
        const bool blk = v == 0x314159265359ull, eos = v == 0x177245385090ull;
        if (blk || eos) {
            const u64 slot = atomicAdd(count, 1ull);
            if (slot < cap) list[slot] = ((i * 8 + b) << 1) | (eos ? 1u : 0u);
        }
    }
}

int find_bzip2_magics(const u8 *d_in, u64 begin, u64 n, std::vector<u64> &entries) {
    entries.clear();
    if (n < begin + 6) return SWC_OK;
    const u64 cap = 1u << 20;
    void *p = nullptr;
    int st = arena_get(2, 256 + cap * 8, &p, 0);
    if (st) return st;
    unsigned long long *count = (unsigned long long *)p;
    u64 *list = (u64 *)((u8 *)p + 256);
    SWC_CUDA_TRY(cudaMemsetAsync(count, 0, 8, 0));
    const u64 threads = n - begin;
    bzip2_magic_scan_kernel<<<(unsigned)((threads + 255) / 256), 256>>>(d_in, begin, n, list, count, cap);
    count_launch();
    SWC_CUDA_TRY(cudaGetLastError());
    unsigned long long h = 0;
    SWC_CUDA_TRY(cudaMemcpy(&h, count, 8, cudaMemcpyDeviceToHost));
    if (h > cap) return SWC_OK;                                       // caller decodes sequentially
    entries.resize(h);
    if (h) SWC_CUDA_TRY(cudaMemcpy(entries.data(), list, h * 8, cudaMemcpyDeviceToHost));
    std::sort(entries.begin(), entries.end());
    return SWC_OK;
}

// gather: unit i's bytes [src_off[i], +len[i]) -> dst[dst_off[i] ...), one warp per unit
__global__ void __launch_bounds__(256) gather_units_kernel(const u8 *src, const u64 *src_off, const u64 *len, u8 *dst, const u64 *dst_off, u64 n) {
    const u64 unit = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (unit >= n) return;
    const u32 lane = threadIdx.x & 31;
    const u8 *s = src + src_off[unit];
    u8 *d = dst + dst_off[unit];
    const u64 L = len[unit];
    for (u64 i = lane; i < L; i += 32) d[i] = s[i];
}

int gather_units(const u8 *src, const u64 *src_off, const u64 *len, u8 *dst, const u64 *dst_off, u64 n, cudaStream_t s) {
    if (!n) return SWC_OK;
    gather_units_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, s>>>(src, src_off, len, dst, dst_off, n);
    count_launch();
    SWC_CUDA_TRY(cudaGetLastError());
    return SWC_OK;
}

// ---------------------------------------------------------------- xxHash32 (seed 0), one thread per buffer
#define XP1 0x9E3779B1u
#define XP2 0x85EBCA77u
#define XP3 0xC2B2AE3Du
#define XP4 0x27D4EB2Fu
#define XP5 0x165667B1u
__device__ __forceinline__ u32 rotl32(u32 v, int s) { return __funnelshift_l(v, v, s); }
__device__ __forceinline__ u32 rd32(const u8 *p) { return (u32)p[0] | (u32)p[1] << 8 | (u32)p[2] << 16 | (u32)p[3] << 24; }

// One WARP per buffer.  The four accumulators of xxHash32 (XxHash32.swift:39-57) are four serial chains over the 16-byte
// stripes — nothing to parallelise there — but a single thread spends most of its time waiting for its own loads.  Here the
// 32 lanes fetch 32 stripes (512 B, coalesced) one chunk ahead into shared memory and lanes 0..3 run one chain each out of
// shared memory: ~1 byte per cycle per buffer instead of ~0.1.
__global__ void __launch_bounds__(128) xxh32_warp_kernel(const u8 *base, const u64 *off, const u64 *len, u32 *result, u64 n) {
    __shared__ u32 stripes[4][2][128];
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u64 unit = (u64)blockIdx.x * 4 + warp;
    if (unit >= n) return;
    const u8 *p = base + (off ? off[unit] : 0);
    const u64 bytes = len[unit] & ~(1ull << 63);
    const u64 nstripes = bytes / 16;
    const bool aligned = (((uintptr_t)p) & 15) == 0;
    auto fetch = [&](u64 s) -> uint4 {                                  // stripe s (caller guarantees s < nstripes)
        const u8 *q = p + s * 16;
        if (aligned) return __ldg((const uint4 *)q);
        return make_uint4(rd32(q), rd32(q + 4), rd32(q + 8), rd32(q + 12));
    };
    u32 acc = lane == 0 ? XP1 + XP2 : lane == 1 ? XP2 : lane == 2 ? 0u : 0u - XP1;
    uint4 next = make_uint4(0, 0, 0, 0);
    if (lane < nstripes) next = fetch(lane);
    int buf = 0;
    for (u64 s0 = 0; s0 < nstripes; s0 += 32, buf ^= 1) {
        u32 *st = stripes[warp][buf];
        *(uint4 *)(st + lane * 4) = next;
        if (s0 + 32 + lane < nstripes) next = fetch(s0 + 32 + lane);   // the next chunk travels while this one is hashed
        __syncwarp();
        const u32 cnt = nstripes - s0 < 32 ? (u32)(nstripes - s0) : 32u;
        if (lane < 4)
            for (u32 k = 0; k < cnt; k++) acc = rotl32(acc + st[k * 4 + lane] * XP2, 13) * XP1;
        __syncwarp();
    }
    const u32 a0 = __shfl_sync(SWC_FULL, acc, 0), a1 = __shfl_sync(SWC_FULL, acc, 1), a2 = __shfl_sync(SWC_FULL, acc, 2), a3 = __shfl_sync(SWC_FULL, acc, 3);
    if (lane == 0) {
        u32 h = bytes < 16 ? XP5 : rotl32(a0, 1) + rotl32(a1, 7) + rotl32(a2, 12) + rotl32(a3, 18);
        h += (u32)bytes;
        u64 i = nstripes * 16;
        for (; bytes - i >= 4; i += 4) h = rotl32(h + rd32(p + i) * XP3, 17) * XP4;
        for (; bytes - i >= 1; i += 1) h = rotl32(h + (u32)p[i] * XP5, 11) * XP1;
        h ^= h >> 15; h *= XP2; h ^= h >> 13; h *= XP3; h ^= h >> 16;
        result[unit] = h;
    }
}

int xxh32_batch(const u8 *base, const u64 *off, const u64 *len, u32 *result, u64 n, cudaStream_t s) {
    if (!n) return SWC_OK;
    xxh32_warp_kernel<<<(unsigned)((n + 3) / 4), 128, 0, s>>>(base, off, len, result, n);
    count_launch();
    SWC_CUDA_TRY(cudaGetLastError());
    return SWC_OK;
}

// ---------------------------------------------------------------- SHA-256, one thread
__constant__ u32 c_k256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
__device__ __forceinline__ u32 rotr32(u32 v, int s) { return __funnelshift_r(v, v, s); }

__device__ void sha256_block(u32 h[8], const u8 *b) {
    u32 w[64];
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = (u32)b[4 * i] << 24 | (u32)b[4 * i + 1] << 16 | (u32)b[4 * i + 2] << 8 | b[4 * i + 3];
#pragma unroll
    for (int i = 16; i < 64; i++) {
        u32 s0 = rotr32(w[i - 15], 7) ^ rotr32(w[i - 15], 18) ^ (w[i - 15] >> 3);
        u32 s1 = rotr32(w[i - 2], 17) ^ rotr32(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    u32 a = h[0], bb = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        u32 t1 = hh + (rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25)) + ((e & f) ^ (~e & g)) + c_k256[i] + w[i];
        u32 t2 = (rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22)) + ((a & bb) ^ (a & c) ^ (bb & c));
        hh = g; g = f; f = e; e = d + t1; d = c; c = bb; bb = a; a = t1 + t2;
    }
    h[0] += a; h[1] += bb; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

__global__ void sha256_kernel(const u8 *p, u64 n, u8 *digest) {
    u32 h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    u64 i = 0;
    for (; i + 64 <= n; i += 64) sha256_block(h, p + i);
    u8 tail[128];
    for (int k = 0; k < 128; k++) tail[k] = 0;
    u64 rem = n - i;
    for (u64 k = 0; k < rem; k++) tail[k] = p[i + k];
    tail[rem] = 0x80;
    int tl = rem + 9 <= 64 ? 64 : 128;
    u64 bits = n * 8;
    for (int k = 0; k < 8; k++) tail[tl - 1 - k] = (u8)(bits >> (8 * k));
    sha256_block(h, tail);
    if (tl == 128) sha256_block(h, tail + 64);
    for (int k = 0; k < 8; k++) { digest[4 * k] = (u8)(h[k] >> 24); digest[4 * k + 1] = (u8)(h[k] >> 16); digest[4 * k + 2] = (u8)(h[k] >> 8); digest[4 * k + 3] = (u8)h[k]; }
}

int sha256(const u8 *d, u64 n, u8 *d_digest, cudaStream_t s) {
    sha256_kernel<<<1, 1, 0, s>>>(d, n, d_digest);
    count_launch();
    SWC_CUDA_TRY(cudaGetLastError());
    return SWC_OK;
}

}  // namespace checks
}  // namespace swc
