// api_deflate.cu — C ABI for Deflate (include/swcgpu.h): batched device call, batched host call, single unit.
#include <cstring>
#include "../../include/swcgpu.h"
#include "host_util.h"
#include "inflate.cuh"

namespace swc {

int to_host_alloc(const void *d, size_t n, uint8_t **out, size_t *out_len) {
    uint8_t *h = (uint8_t *)swc_alloc(n);
    if (!h) return SWC_ERR_OUTPUT_OVERFLOW;
    if (n) {
        cudaError_t e = cudaMemcpy(h, d, n, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { swc_free(h); return cuda_fail(e, "cudaMemcpy D2H"); }
    }
    *out = h;
    *out_len = n;
    return SWC_OK;
}

static int deflate_batch_impl(const uint8_t *in_base, const uint64_t *in_off, const uint64_t *in_len, const uint8_t *start_bits,
                              uint8_t *out_base, const uint64_t *out_off, const uint64_t *out_cap, uint64_t out_total,
                              uint64_t *out_len, uint64_t *consumed_bits, int32_t *status, uint64_t n,
                              void *scratch, size_t scratch_bytes, cudaStream_t stream) {
    if (n == 0) return SWC_OK;
    if (!in_base || !in_off || !in_len || !out_base || !out_off || !out_cap || !out_len || !consumed_bits || !status)
        return SWC_ERR_INVALID_ARG;
    const size_t need = inflate::scratch_bytes(n, out_total);
    if (!scratch) {
        int st = scratch_get(need, &scratch, stream);
        if (st) return st;
    } else if (scratch_bytes < need) {
        return SWC_ERR_INVALID_ARG;
    }
    inflate::BatchArgs a;
    a.in_base = in_base; a.in_off = in_off; a.in_len = in_len; a.start_bits = start_bits;
    a.out_base = out_base; a.out_off = out_off; a.out_cap = out_cap;
    a.out_len = out_len; a.consumed_bits = consumed_bits; a.status = status; a.n = n;
    a.ticket = (unsigned long long *)scratch;                       // 8 bytes (256 reserved)
    a.rec_count = (u32 *)((u8 *)scratch + 256);                     // n entries
    a.rec_base = (u32 *)((u8 *)scratch + 256 + ((n * 4 + 255) & ~(size_t)255));
    return inflate::launch(a, stream);
}

// Decode one Deflate stream that already sits on the device; grows the output buffer until it fits.
int deflate_unit_device(const u8 *d_in, size_t in_len, size_t start_bit_abs, UnitResult &r, size_t hint) {
    size_t start_byte = start_bit_abs >> 3;
    if (start_byte > in_len) start_byte = in_len;
    u8 sb = (u8)(start_bit_abs & 7);
    size_t cap = hint ? hint : (in_len - start_byte) * 4 + 65536;
    DevBuf meta;
    int st = meta.alloc(256);
    if (st) return st;
    for (int attempt = 0; attempt < 3; attempt++) {
        cap = round16(cap);
        if ((st = r.out.alloc(cap))) return st;
        // meta layout: in_off, in_len, out_off, out_cap, out_len, consumed (u64 each) | status (i32) | start_bits (u8)
        u64 h[6] = {start_byte, in_len - start_byte, 0, cap, 0, 0};
        SWC_CUDA_TRY(cudaMemcpy(meta.p, h, sizeof(h), cudaMemcpyHostToDevice));
        SWC_CUDA_TRY(cudaMemcpy((u8 *)meta.p + 64, &sb, 1, cudaMemcpyHostToDevice));
        u64 *m = meta.as<u64>();
        st = deflate_batch_impl(d_in, m + 0, m + 1, (u8 *)meta.p + 64, r.out.as<u8>(), m + 2, m + 3, cap, m + 4, m + 5,
                                (int32_t *)((u8 *)meta.p + 48), 1, nullptr, 0, 0);
        if (st) return st;
        SWC_CUDA_TRY(cudaStreamSynchronize(0));
        u64 res[7];
        SWC_CUDA_TRY(cudaMemcpy(res, meta.p, 56, cudaMemcpyDeviceToHost));
        r.out_len = (size_t)res[4];
        r.consumed = (size_t)res[5];
        int32_t st32; memcpy(&st32, &res[6], 4);
        r.status = st32;
        if (r.status != SWC_ERR_OUTPUT_OVERFLOW) break;
        cap = r.out_len;          // exact size reported by the counting pass
    }
    if (r.status == SWC_INTERNAL_NEEDS_SLOW) r.status = SWC_ERR_UNSUPPORTED;
    return SWC_OK;
}

}  // namespace swc

using namespace swc;

extern "C" {

size_t swc_deflate_batch_scratch_bytes(uint64_t n, uint64_t out_capacity_total) {
    return inflate::scratch_bytes(n, out_capacity_total);
}

int32_t swc_deflate_decompress_batch(const uint8_t *in_base, const uint64_t *in_off, const uint64_t *in_len,
                                     const uint8_t *start_bits,
                                     uint8_t *out_base, const uint64_t *out_off, const uint64_t *out_cap,
                                     uint64_t out_capacity_total,
                                     uint64_t *out_len, uint64_t *consumed_bits, int32_t *status,
                                     uint64_t n, void *scratch, size_t scratch_bytes, void *cuda_stream) {
    if (ensure_device()) return SWC_ERR_NO_DEVICE;
    return deflate_batch_impl(in_base, in_off, in_len, start_bits, out_base, out_off, out_cap, out_capacity_total,
                              out_len, consumed_bits, status, n, scratch, scratch_bytes, (cudaStream_t)cuda_stream);
}

// Host-buffer batch: H2D of the compressed bytes + tables, decode, D2H of the decoded bytes + results.
int32_t swc_deflate_decompress_batch_host(const uint8_t *in_base, const uint64_t *in_off, const uint64_t *in_len,
                                          uint64_t in_total,
                                          uint8_t *out_base, const uint64_t *out_off, const uint64_t *out_cap,
                                          uint64_t out_total,
                                          uint64_t *out_len, uint64_t *consumed_bits, int32_t *status, uint64_t n) {
    if (ensure_device()) return SWC_ERR_NO_DEVICE;
    if (n == 0) return SWC_OK;
    DevBuf d_in, d_out, d_meta;
    int st;
    if ((st = d_in.alloc(round16(in_total) + 16))) return st;
    if ((st = d_out.alloc(round16(out_total)))) return st;
    const size_t tb = n * 8;
    if ((st = d_meta.alloc(tb * 6 + n * 4))) return st;
    u8 *m = d_meta.as<u8>();
    cudaStream_t s = 0;
    SWC_CUDA_TRY(cudaMemcpyAsync(d_in.p, in_base, in_total, cudaMemcpyHostToDevice, s));
    SWC_CUDA_TRY(cudaMemcpyAsync(m + 0 * tb, in_off, tb, cudaMemcpyHostToDevice, s));
    SWC_CUDA_TRY(cudaMemcpyAsync(m + 1 * tb, in_len, tb, cudaMemcpyHostToDevice, s));
    SWC_CUDA_TRY(cudaMemcpyAsync(m + 2 * tb, out_off, tb, cudaMemcpyHostToDevice, s));
    SWC_CUDA_TRY(cudaMemcpyAsync(m + 3 * tb, out_cap, tb, cudaMemcpyHostToDevice, s));
    st = deflate_batch_impl(d_in.as<u8>(), (u64 *)(m + 0 * tb), (u64 *)(m + 1 * tb), nullptr, d_out.as<u8>(),
                            (u64 *)(m + 2 * tb), (u64 *)(m + 3 * tb), out_total, (u64 *)(m + 4 * tb), (u64 *)(m + 5 * tb),
                            (int32_t *)(m + 6 * tb), n, nullptr, 0, s);
    if (st) return st;
    SWC_CUDA_TRY(cudaMemcpyAsync(out_base, d_out.p, out_total, cudaMemcpyDeviceToHost, s));
    SWC_CUDA_TRY(cudaMemcpyAsync(out_len, m + 4 * tb, tb, cudaMemcpyDeviceToHost, s));
    SWC_CUDA_TRY(cudaMemcpyAsync(consumed_bits, m + 5 * tb, tb, cudaMemcpyDeviceToHost, s));
    SWC_CUDA_TRY(cudaMemcpyAsync(status, m + 6 * tb, n * 4, cudaMemcpyDeviceToHost, s));
    SWC_CUDA_TRY(cudaStreamSynchronize(s));
    for (uint64_t i = 0; i < n; i++)
        if (status[i] == SWC_INTERNAL_NEEDS_SLOW) status[i] = SWC_ERR_UNSUPPORTED;
    return SWC_OK;
}

int32_t swc_deflate_decompress(const uint8_t *in, size_t in_len, size_t start_bit,
                               uint8_t **out, size_t *out_len, size_t *consumed_bits) {
    if (!out || !out_len) return SWC_ERR_INVALID_ARG;
    *out = nullptr; *out_len = 0;
    if (consumed_bits) *consumed_bits = 0;
    if (ensure_device()) return SWC_ERR_NO_DEVICE;
    DevBuf d_in;
    int st = d_in.alloc(round16(in_len) + 16);
    if (st) return st;
    if (in_len) SWC_CUDA_TRY(cudaMemcpy(d_in.p, in, in_len, cudaMemcpyHostToDevice));
    UnitResult r;
    if ((st = deflate_unit_device(d_in.as<u8>(), in_len, start_bit, r))) return st;
    if (consumed_bits) *consumed_bits = r.consumed;
    if (r.status != SWC_OK) return r.status;
    return to_host_alloc(r.out.p, r.out_len, out, out_len);
}

}  // extern "C"
