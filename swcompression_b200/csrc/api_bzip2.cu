// api_bzip2.cu — C ABI for BZip2 (include/swcgpu.h): batched device call, single stream, multi-stream.
#include <cstring>
#include <vector>
#include "../../include/swcgpu.h"
#include "host_util.h"
#include "bzip2.cuh"
#include "checks.cuh"

using namespace swc;

namespace {

int bzip2_batch_impl(const uint8_t *in_base, const uint64_t *in_off, const uint64_t *in_len,
                     uint8_t *out_base, const uint64_t *out_off, const uint64_t *out_cap,
                     uint64_t *out_len, uint64_t *consumed_bits, int32_t *status, uint64_t n, cudaStream_t stream) {
    if (n == 0) return SWC_OK;
    if (!in_base || !in_off || !in_len || !out_base || !out_off || !out_cap || !out_len || !consumed_bits || !status)
        return SWC_ERR_INVALID_ARG;
    // per-unit scratch sizes depend on the capacities: fetch them (n x 8 bytes), lay the scratch out on the host
    std::vector<uint64_t> caps(n), soff(n);
    SWC_CUDA_TRY(cudaMemcpyAsync(caps.data(), out_cap, n * 8, cudaMemcpyDeviceToHost, stream));
    SWC_CUDA_TRY(cudaStreamSynchronize(stream));
    size_t total = (n * 8 + 255) & ~(size_t)255;
    for (uint64_t i = 0; i < n; i++) { soff[i] = total; total += bzip2::scratch_per_unit(caps[i]); }
    void *scratch = nullptr;
    int st = scratch_get(total, &scratch, stream);
    if (st) return st;
    SWC_CUDA_TRY(cudaMemcpyAsync(scratch, soff.data(), n * 8, cudaMemcpyHostToDevice, stream));
    SWC_CUDA_TRY(cudaStreamSynchronize(stream));     // soff lives on this stack frame
    bzip2::Args a;
    a.in_base = in_base; a.in_off = in_off; a.in_len = in_len;
    a.out_base = out_base; a.out_off = out_off; a.out_cap = out_cap;
    a.out_len = out_len; a.consumed_bits = consumed_bits; a.status = status; a.n = n;
    a.scratch = (u8 *)scratch; a.scr_off = (const u64 *)scratch;
    return bzip2::launch(a, stream);
}

// 32 bits at an arbitrary bit position of the host copy, MSB first (BZip2's reader order)
uint32_t bits32_at(const uint8_t *in, size_t in_len, uint64_t bitpos) {
    uint64_t w = 0;
    const size_t b0 = (size_t)(bitpos >> 3);
    for (int k = 0; k < 5; k++) w = (w << 8) | (b0 + k < in_len ? in[b0 + k] : 0);
    return (uint32_t)(w >> (8 - (bitpos & 7)));
}

// Block-parallel decode of ONE stream.  The blocks of a stream are independent once their starts are known, and the starts
// are marked by a 48-bit magic — at any bit offset.  All magics are found by a device scan, every block candidate is decoded
// as its own unit of one batched launch (block mode of the kernel), and the reference's sequential walk (BZip2.swift:66-95)
// is kept as the validator: block k is accepted iff it decoded cleanly (including its CRC) and ended exactly at the next
// magic; the walk must end on the end-of-stream magic with the right combined CRC.  Anything else (a look-alike magic, an
// overflowing block, damage) leaves *done == false and the caller decodes the stream the sequential way, which also
// produces the reference's error.  `in` is the host copy of the archive, d_in the device copy.
int bzip2_stream_by_blocks(const uint8_t *in, const u8 *d_in, size_t in_len, size_t start, UnitResult &r, bool *done) {
    *done = false;
    if (in_len < start + 14) return SWC_OK;
    if (in[start] != 'B' || in[start + 1] != 'Z' || in[start + 2] != 'h' || in[start + 3] < '1' || in[start + 3] > '9') return SWC_OK;
    const size_t level = (size_t)(in[start + 3] - '0');
    std::vector<u64> mags;
    int st = checks::find_bzip2_magics(d_in, start + 4, in_len, mags);
    if (st) return st;
    // blocks from the first magic (which must sit right behind the header) up to the first end-of-stream magic
    std::vector<u64> pos;
    u64 end_pos = 0; bool have_end = false;
    for (u64 e : mags) {
        if (e & 1) { end_pos = e >> 1; have_end = true; break; }
        pos.push_back(e >> 1);
    }
    if (!have_end || pos.size() < 2 || pos[0] != (u64)(start + 4) * 8) return SWC_OK;
    const size_t nb = pos.size();
    const size_t cap = round16(level * 100000 * 2 + 65536);          // RLE1 can expand a block further: then the sequential path takes over
    std::vector<u64> h(nb * 4);
    std::vector<u8> h_sb(nb);
    for (size_t k = 0; k < nb; k++) {
        h[k] = pos[k] >> 3; h[nb + k] = in_len - (pos[k] >> 3); h[2 * nb + k] = k * cap; h[3 * nb + k] = cap;
        h_sb[k] = (u8)(pos[k] & 7);
    }
    // the block-parallel attempt needs nb x (cap + ~6 x cap of scratch); when that does not fit, the sequential decoder (which
    // holds one block at a time) takes over instead of failing the call
    {
        size_t free_b = 0, total_b = 0;
        SWC_CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        const size_t need = nb * (cap + bzip2::scratch_per_unit(cap)) + ((size_t)64 << 20);
        if (need > free_b / 2) return SWC_OK;
    }
    DevBuf d_out, d_meta;
    if (d_out.alloc(nb * cap + 64)) return SWC_OK;
    if (d_meta.alloc(nb * 8 * 7 + nb * 4 + nb + 64)) return SWC_OK;
    u64 *m = d_meta.as<u64>();
    int32_t *d_status = (int32_t *)(m + 7 * nb);
    u8 *d_sb = (u8 *)(d_status + nb);
    SWC_CUDA_TRY(cudaMemcpy(m, h.data(), nb * 32, cudaMemcpyHostToDevice));
    SWC_CUDA_TRY(cudaMemcpy(d_sb, h_sb.data(), nb, cudaMemcpyHostToDevice));
    {
        size_t total = (nb * 8 + 255) & ~(size_t)255;
        std::vector<u64> soff(nb);
        for (size_t k = 0; k < nb; k++) { soff[k] = total; total += bzip2::scratch_per_unit(cap); }
        void *scratch = nullptr;
        if (scratch_get(total, &scratch, 0)) return SWC_OK;
        SWC_CUDA_TRY(cudaMemcpy(scratch, soff.data(), nb * 8, cudaMemcpyHostToDevice));
        bzip2::Args a;
        a.in_base = d_in; a.in_off = m; a.in_len = m + nb;
        a.out_base = d_out.as<u8>(); a.out_off = m + 2 * nb; a.out_cap = m + 3 * nb;
        a.out_len = m + 4 * nb; a.consumed_bits = m + 5 * nb; a.status = d_status; a.n = nb;
        a.scratch = (u8 *)scratch; a.scr_off = (const u64 *)scratch;
        a.start_bits = d_sb; a.block_mode = 1;
        if ((st = bzip2::launch(a, 0))) return st;
    }
    std::vector<u64> r_len(nb), r_used(nb);
    std::vector<int32_t> r_st(nb);
    SWC_CUDA_TRY(cudaMemcpy(r_len.data(), m + 4 * nb, nb * 8, cudaMemcpyDeviceToHost));
    SWC_CUDA_TRY(cudaMemcpy(r_used.data(), m + 5 * nb, nb * 8, cudaMemcpyDeviceToHost));
    SWC_CUDA_TRY(cudaMemcpy(r_st.data(), d_status, nb * 4, cudaMemcpyDeviceToHost));
    // the in-order walk: every block must end where the next magic starts, the last one at the end-of-stream magic
    uint32_t total_crc = 0;
    std::vector<u64> g_dst(nb);
    u64 out_total = 0;
    for (size_t k = 0; k < nb; k++) {
        const u64 ends_at = (pos[k] & ~7ull) + r_used[k];            // consumed is counted from the unit's first byte
        const u64 next = k + 1 < nb ? pos[k + 1] : end_pos;
        if (r_st[k] != SWC_OK || ends_at != next) return SWC_OK;
        const uint32_t block_crc = bits32_at(in, in_len, pos[k] + 48);
        total_crc = ((total_crc << 1) | (total_crc >> 31)) ^ block_crc;  // BZip2.swift:83-84
        g_dst[k] = out_total;
        out_total += r_len[k];
    }
    if (end_pos + 80 > (u64)in_len * 8) return SWC_OK;
    if (bits32_at(in, in_len, end_pos + 48) != total_crc) return SWC_OK;   // the sequential path reports wrongCRC with its payload
    // contiguous result
    if ((st = r.out.alloc(round16((size_t)out_total) + 64))) return st;
    DevBuf d_g;
    if ((st = d_g.alloc(nb * 8 + 64))) return st;
    SWC_CUDA_TRY(cudaMemcpy(d_g.p, g_dst.data(), nb * 8, cudaMemcpyHostToDevice));
    if ((st = checks::gather_units(d_out.as<u8>(), m + 2 * nb, m + 4 * nb, r.out.as<u8>(), d_g.as<u64>(), nb, 0))) return st;
    SWC_CUDA_TRY(cudaStreamSynchronize(0));
    r.out_len = (size_t)out_total;
    r.consumed = (size_t)(end_pos + 80 - (u64)start * 8);
    r.status = SWC_OK;
    *done = true;
    return SWC_OK;
}

// one stream starting at byte `start` of d_in; output grows until it fits.  `in` (optional) = host copy of the archive:
// with it, multi-block streams are first tried block-parallel.
int bzip2_unit_device(const u8 *d_in, size_t in_len, size_t start, UnitResult &r, const uint8_t *in = nullptr) {
    if (in) {
        bool done = false;
        int bst = bzip2_stream_by_blocks(in, d_in, in_len, start, r, &done);
        if (bst) return bst;
        if (done) return SWC_OK;
    }
    size_t cap = (in_len - start) * 16 + (1u << 20);
    DevBuf meta;
    int st = meta.alloc(256);
    if (st) return st;
    for (int attempt = 0; attempt < 8; attempt++) {
        cap = round16(cap);
        if ((st = r.out.alloc(cap))) return st;
        u64 h[6] = {start, in_len - start, 0, cap, 0, 0};
        SWC_CUDA_TRY(cudaMemcpy(meta.p, h, sizeof(h), cudaMemcpyHostToDevice));
        u64 *m = meta.as<u64>();
        if ((st = bzip2_batch_impl(d_in, m + 0, m + 1, r.out.as<u8>(), m + 2, m + 3, m + 4, m + 5, (int32_t *)((u8 *)meta.p + 48), 1, 0))) return st;
        SWC_CUDA_TRY(cudaStreamSynchronize(0));
        u64 res[7];
        SWC_CUDA_TRY(cudaMemcpy(res, meta.p, 56, cudaMemcpyDeviceToHost));
        r.out_len = (size_t)res[4];
        r.consumed = (size_t)res[5];
        int32_t st32; memcpy(&st32, &res[6], 4);
        r.status = st32;
        if (r.status != SWC_ERR_OUTPUT_OVERFLOW || cap >= ((size_t)1 << 32)) break;
        cap *= 4;                     // BZip2 cannot report the final size before it has decoded everything
    }
    return SWC_OK;
}

int upload(DevBuf &d, const uint8_t *in, size_t n) {
    int st = d.alloc(round16(n) + 256);
    if (st) return st;
    { int cst = copy_pageable(d.p, in, n, true); if (cst) return cst; }
    return SWC_OK;
}

}  // namespace

extern "C" {

int32_t swc_bzip2_decompress_batch(const uint8_t *in_base, const uint64_t *in_off, const uint64_t *in_len,
                                   uint8_t *out_base, const uint64_t *out_off, const uint64_t *out_cap,
                                   uint64_t *out_len, uint64_t *consumed_bits, int32_t *status,
                                   uint64_t n, void *cuda_stream) {
    if (ensure_device()) return SWC_ERR_NO_DEVICE;
    ApiLock api_lock;
    return bzip2_batch_impl(in_base, in_off, in_len, out_base, out_off, out_cap, out_len, consumed_bits, status, n, (cudaStream_t)cuda_stream);
}

// BZip2.decompress(data:) BZip2.swift:22-26 (+ reader form :50)
int32_t swc_bzip2_decompress(const uint8_t *in, size_t in_len, size_t start_bit,
                             uint8_t **out, size_t *out_len, size_t *consumed_bits) {
    if (!out || !out_len) return SWC_ERR_INVALID_ARG;
    *out = nullptr; *out_len = 0;
    if (consumed_bits) *consumed_bits = 0;
    if (start_bit & 7) return SWC_ERR_UNSUPPORTED;      // the reference's byte reads require an aligned reader (BZip2.swift:59)
    if (ensure_device()) return SWC_ERR_NO_DEVICE;
    ApiLock api_lock;
    DevBuf d_in;
    int st = upload(d_in, in, in_len);
    if (st) return st;
    UnitResult r;
    if ((st = bzip2_unit_device(d_in.as<u8>(), in_len, start_bit >> 3, r, in))) return st;
    if (consumed_bits) *consumed_bits = r.consumed;
    if (r.status != SWC_OK && r.status != SWC_BZIP2_WRONG_CRC) return r.status;
    if ((st = to_host_alloc(r.out.p, r.out_len, out, out_len))) return st;
    return r.status;
}

// BZip2.multiDecompress(data:) BZip2.swift:40-48
int32_t swc_bzip2_multi_decompress(const uint8_t *in, size_t in_len,
                                   uint8_t **out, size_t *out_len, size_t **stream_ends, size_t *n_streams) {
    if (!out || !out_len || !stream_ends || !n_streams) return SWC_ERR_INVALID_ARG;
    *out = nullptr; *out_len = 0; *stream_ends = nullptr; *n_streams = 0;
    if (ensure_device()) return SWC_ERR_NO_DEVICE;
    ApiLock api_lock;
    DevBuf d_in;
    int st = upload(d_in, in, in_len);
    if (st) return st;
    std::vector<uint8_t> o;
    std::vector<size_t> ends;
    size_t off = 0;
    int result = SWC_OK;
    // ---- fast path: discover the streams up front and decode them as ONE batch ------------------------------------
    // A stream starts byte-aligned with 'B' 'Z' 'h' '1'..'9' followed by a block magic (or the end magic of an empty
    // stream).  Candidates found by scanning for that 10-byte signature are decoded in parallel; a candidate is accepted
    // only if the previous stream ended exactly there (consumed bytes == distance to the next candidate).  The first
    // stream that does not validate — or fails — hands over to the sequential loop below, which IS the reference's
    // order of events (BZip2.swift:40-48), so results and errors are the same with or without the fast path.
    {
        std::vector<size_t> cand;
        static const uint8_t blk[6] = {0x31, 0x41, 0x59, 0x26, 0x53, 0x59}, eos[6] = {0x17, 0x72, 0x45, 0x38, 0x50, 0x90};
        for (size_t i = 0; i + 10 <= in_len; i++) {
            if (in[i] != 'B' || in[i + 1] != 'Z' || in[i + 2] != 'h' || in[i + 3] < '1' || in[i + 3] > '9') continue;
            if (memcmp(in + i + 4, blk, 6) == 0 || memcmp(in + i + 4, eos, 6) == 0) cand.push_back(i);
        }
        if (cand.size() >= 2 && cand[0] == 0) {
            const size_t n = cand.size();
            std::vector<uint64_t> h_off(n), h_len(n), h_ooff(n), h_cap(n), r_len(n), r_used(n);
            std::vector<int32_t> r_st(n);
            // groups bounded by a scratch budget (the BWT working set is ~6x the output capacity of a unit)
            const size_t budget = (size_t)24 << 30;
            size_t k0 = 0;
            bool handed_over = false;
            while (k0 < n && !handed_over) {
                size_t k1 = k0, need = 0, out_total = 0;
                while (k1 < n) {
                    const size_t len = (k1 + 1 < n ? cand[k1 + 1] : in_len) - cand[k1];
                    const size_t cap = round16(len * 12 + (1u << 20));
                    const size_t add = bzip2::scratch_per_unit(cap) + cap;
                    if (k1 > k0 && need + add > budget) break;
                    h_off[k1] = cand[k1]; h_len[k1] = len; h_ooff[k1] = out_total; h_cap[k1] = cap;
                    need += add; out_total += cap; k1++;
                }
                const size_t g = k1 - k0;
                DevBuf d_out, d_meta;
                if ((st = d_out.alloc(out_total + 64))) return st;
                if ((st = d_meta.alloc(g * 8 * 6 + g * 4 + 64))) return st;
                u64 *m = d_meta.as<u64>();
                SWC_CUDA_TRY(cudaMemcpy(m + 0 * g, h_off.data() + k0, g * 8, cudaMemcpyHostToDevice));
                SWC_CUDA_TRY(cudaMemcpy(m + 1 * g, h_len.data() + k0, g * 8, cudaMemcpyHostToDevice));
                SWC_CUDA_TRY(cudaMemcpy(m + 2 * g, h_ooff.data() + k0, g * 8, cudaMemcpyHostToDevice));
                SWC_CUDA_TRY(cudaMemcpy(m + 3 * g, h_cap.data() + k0, g * 8, cudaMemcpyHostToDevice));
                if ((st = bzip2_batch_impl(d_in.as<u8>(), m + 0 * g, m + 1 * g, d_out.as<u8>(), m + 2 * g, m + 3 * g, m + 4 * g, m + 5 * g,
                                           (int32_t *)(m + 6 * g), g, 0))) return st;
                SWC_CUDA_TRY(cudaStreamSynchronize(0));
                SWC_CUDA_TRY(cudaMemcpy(r_len.data() + k0, m + 4 * g, g * 8, cudaMemcpyDeviceToHost));
                SWC_CUDA_TRY(cudaMemcpy(r_used.data() + k0, m + 5 * g, g * 8, cudaMemcpyDeviceToHost));
                SWC_CUDA_TRY(cudaMemcpy(r_st.data() + k0, m + 6 * g, g * 4, cudaMemcpyDeviceToHost));
                for (size_t k = k0; k < k1; k++) {
                    const bool ok = r_st[k] == SWC_OK && (r_used[k] + 7) / 8 == h_len[k];
                    if (!ok) { off = cand[k]; handed_over = true; break; }        // sequential loop re-decodes this stream
                    const size_t base = o.size();
                    o.resize(base + r_len[k]);
                    if (r_len[k]) SWC_CUDA_TRY(cudaMemcpy(o.data() + base, d_out.as<u8>() + h_ooff[k], r_len[k], cudaMemcpyDeviceToHost));
                    ends.push_back(o.size());
                    off = cand[k] + h_len[k];
                }
                k0 = k1;
            }
        }
    }
    while (off < in_len) {                               // !reader.isFinished
        UnitResult r;
        if ((st = bzip2_unit_device(d_in.as<u8>(), in_len, off, r, in))) return st;
        if (r.status != SWC_OK && r.status != SWC_BZIP2_WRONG_CRC) return r.status;
        size_t base = o.size();
        if (r.status == SWC_BZIP2_WRONG_CRC) { o.clear(); ends.clear(); base = 0; }   // payload = the failing archive only
        o.resize(base + r.out_len);
        if (r.out_len) SWC_CUDA_TRY(cudaMemcpy(o.data() + base, r.out.p, r.out_len, cudaMemcpyDeviceToHost));
        ends.push_back(o.size());
        if (r.status) { result = r.status; break; }
        off += (r.consumed + 7) / 8;                     // reader.align()
    }
    uint8_t *h = (uint8_t *)swc_alloc(o.size());
    if (!h) return SWC_ERR_OUTPUT_OVERFLOW;
    if (!o.empty()) memcpy(h, o.data(), o.size());
    *out = h; *out_len = o.size();
    *stream_ends = (size_t *)swc_alloc(sizeof(size_t) * (ends.size() + 1));
    for (size_t i = 0; i < ends.size(); i++) (*stream_ends)[i] = ends[i];
    *n_streams = ends.size();
    return result;
}

}  // extern "C"
