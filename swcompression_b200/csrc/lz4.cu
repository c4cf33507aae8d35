// lz4.cu — batched LZ4 block decode for sm_100a.  Replaces LZ4.process(block:_:) (reference
// Sources/LZ4/LZ4.swift:332-413) and the per-block part of LZ4.process(frame:) (:278-318).
//
// ONE WARP PER UNIT.  A unit is a chain of one or more blocks that are decoded in order into one contiguous output
// region (chains model frames with dependent blocks, LZ4.swift:307-313; independent blocks are one-block units).
// All 32 lanes parse the sequence stream in lock-step from a 32-byte register window (one coalesced load per
// sequence in the common case, bytes exchanged with warp shuffles), then copy literals and matches cooperatively.
#include "common.cuh"
#include "lz4.cuh"

namespace swc {
namespace lz4 {

struct Window {
    const u8 *in;
    u64 n;        // block length
    u64 base;     // offset of window byte 0
    u32 w;        // this lane's byte
    __device__ __forceinline__ void load(u64 at) {
        base = at;
        w = (at + lane_id() < n) ? in[at + lane_id()] : 0;
    }
    // make [at, at+need) visible (need <= 32)
    __device__ __forceinline__ void ensure(u64 at, u32 need) {
        if (at < base || at + need > base + 32) load(at);
    }
    __device__ __forceinline__ u32 get(u64 at) const { return __shfl_sync(SWC_FULL, w, (int)(at - base)); }
};

// warp-cooperative forward copy of n bytes, regions do not overlap (or dst > src + 32*16)
__device__ __forceinline__ void warp_copy(u8 *dst, const u8 *src, u64 n) {
    const u32 lane = lane_id();
    if (n >= 256 && (((uintptr_t)dst ^ (uintptr_t)src) & 15) == 0) {
        u64 head = (16 - ((uintptr_t)dst & 15)) & 15;
        if (lane < head) dst[lane] = src[lane];
        u64 body = (n - head) >> 4;
        const uint4 *s4 = (const uint4 *)(src + head);
        uint4 *d4 = (uint4 *)(dst + head);
        for (u64 i = lane; i < body; i += 32) d4[i] = s4[i];
        u64 done = head + (body << 4);
        if (done + lane < n) dst[done + lane] = src[done + lane];
        return;
    }
    for (u64 i = lane; i < n; i += 32) dst[i] = src[i];
}

// length-extension bytes (LZ4.swift:347-363 / :386-401). Returns false on truncation. `ip` ends after the last byte.
__device__ __forceinline__ bool read_extension(Window &win, u64 &ip, u64 &value) {
    const u32 lane = lane_id();
    for (;;) {
        if (ip >= win.n) return false;
        win.ensure(ip, 1);
        u32 pos = (u32)(ip - win.base);
        u32 m = __ballot_sync(SWC_FULL, win.w != 255) & (0xFFFFFFFFu << pos);
        if (m) {
            u32 j = __ffs(m) - 1;
            if (win.base + j >= win.n) return false;        // ran off the block inside the 255-run
            value += 255ull * (j - pos) + __shfl_sync(SWC_FULL, win.w, j);
            ip = win.base + j + 1;
            return true;
        }
        value += 255ull * (32 - pos);
        ip = win.base + 32;
        (void)lane;
    }
}

__global__ void __launch_bounds__(256)
lz4_block_kernel(Args a) {
    const u64 unit = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (unit >= a.n) return;
    const u32 lane = lane_id();
    u8 *out = a.out_base + a.out_off[unit];
    const u64 cap = a.out_cap[unit];
    const u64 b0 = a.first_blk ? a.first_blk[unit] : unit;
    const u64 nb = a.n_blk ? a.n_blk[unit] : 1;
    const i64 dict_len = (i64)a.dict_len;
    const u8 *dict_end = a.dict + a.dict_len;       // dict bytes sit at virtual positions [-dict_len, 0)
    u64 op = 0, fail_blk = 0;
    int status = SWC_OK;

    for (u64 b = 0; b < nb && status == SWC_OK; b++) {
        fail_blk = b;
        const u64 blen_raw = a.blk_len[b0 + b];
        const bool stored = (blen_raw >> 63) != 0;
        const u64 n = blen_raw & ~(1ull << 63);
        const u8 *in = a.in_base + a.blk_off[b0 + b];
        if (stored) {                                                   // LZ4.swift:314-316
            if (op + n <= cap) warp_copy(out + op, in, n);
            op += n;
            __syncwarp();
            continue;
        }
        // lowest addressable virtual position for this block (LZ4.swift:305-313)
        const i64 lo = (op == 0) ? -dict_len : (i64)(op > 65536 ? op - 65536 : 0);   // `out.isEmpty` rule, :307
        Window win; win.in = in; win.n = n; win.base = 0; win.w = 0;
        u64 ip = 0;
        u64 seq = 0;
        i64 last_match_start = -1;
        bool have_match = false;
        if (n > 0) win.load(0);
        for (;;) {
            seq++;
            if (ip >= n) { status = SWC_DATA_TRUNCATED; break; }                         // :343
            win.ensure(ip, 1);
            const u32 token = win.get(ip);
            ip++;
            u64 lit = token >> 4;
            if (lit == 15 && !read_extension(win, ip, lit)) { status = SWC_DATA_TRUNCATED; break; }
            if (n - ip < lit) { status = SWC_DATA_TRUNCATED; break; }                    // :364
            if (op + lit <= cap) warp_copy(out + op, in + ip, lit);
            op += lit;
            ip += lit;
            if (ip == n) {                                                               // :369-377
                if (!(lit >= 5 || seq == 1)) status = SWC_DATA_CORRUPTED;
                else if (have_match && (i64)op - last_match_start < 12) status = SWC_DATA_CORRUPTED;
                break;
            }
            if (n - ip < 2) { status = SWC_DATA_TRUNCATED; break; }                      // :379
            win.ensure(ip, 2);
            const u64 offset = win.get(ip) | (win.get(ip + 1) << 8);
            ip += 2;
            if (offset == 0 || (i64)offset > (i64)op - lo) { status = SWC_DATA_CORRUPTED; break; }   // :383
            u64 mlen = 4 + (token & 15);
            if (mlen == 19 && !read_extension(win, ip, mlen)) { status = SWC_DATA_TRUNCATED; break; }
            last_match_start = (i64)op; have_match = true;                               // :405
            if (op + mlen <= cap) {
                __syncwarp();     // literals just written by other lanes may be match sources
                const i64 src = (i64)op - (i64)offset;
                if (src >= 0 && offset >= mlen && offset >= 16 * 32) {
                    warp_copy(out + op, out + src, mlen);
                } else {
                    for (u64 i = lane; i < mlen; i += 32) {
                        const i64 s = src + (i64)(offset >= mlen ? i : i % offset);
                        out[op + i] = s >= 0 ? out[s] : dict_end[s];
                    }
                }
                __syncwarp();
            }
            op += mlen;
        }
    }
    if (lane == 0) {
        if (status == SWC_OK && op > cap) status = SWC_ERR_OUTPUT_OVERFLOW;
        // chains report WHICH block failed so the host can order errors like the reference's sequential loop
        a.out_len[unit] = (a.n_blk && status != SWC_OK && status != SWC_ERR_OUTPUT_OVERFLOW) ? fail_blk : op;
        a.status[unit] = status;
    }
}

int launch(const Args &a, cudaStream_t stream) {
    if (a.n == 0) return SWC_OK;
    const u64 g = (a.n * 32 + 255) / 256;
    lz4_block_kernel<<<(unsigned)g, 256, 0, stream>>>(a);
    count_launch();
    SWC_CUDA_TRY(cudaGetLastError());
    return SWC_OK;
}

}  // namespace lz4
}  // namespace swc
