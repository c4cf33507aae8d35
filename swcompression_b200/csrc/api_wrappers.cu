// api_wrappers.cu — GzipArchive / ZlibArchive framing around the device Deflate decoder, and the checksum C ABI.
// Reference: Sources/GZip/GzipArchive.swift:38-100, GzipHeader.swift:68-199, Sources/Zlib/ZlibArchive.swift:25-42,
// ZlibHeader.swift:47-93.  Headers/trailers (tens of bytes) are parsed on the host; Deflate and CRC-32 / Adler-32 over
// the payload run on the device.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../include/swcgpu.h"
#include "host_util.h"
#include "checks.cuh"

using namespace swc;

namespace swc {
namespace checks {

int check_device(Kind k, const u8 *d, u64 n, u64 *value) {
    DevBuf res, part;
    int st;
    {   // result slot + chunk partials live in a grow-only arena: per-block checks of big archives do no cudaMalloc
        void *a = nullptr;
        if ((st = arena_get(3, 64 + partial_bytes(n), &a, 0))) return st;
        res.borrow(a, 32);
        part.borrow((u8 *)a + 64, partial_bytes(n));
    }
    if (k == XXH32) {
        SWC_CUDA_TRY(cudaMemcpy(res.p, &n, 8, cudaMemcpyHostToDevice));
        if ((st = xxh32_batch(d, nullptr, res.as<u64>(), (u32 *)(res.as<u8>() + 8), 1, 0))) return st;
        u32 v = 0;
        SWC_CUDA_TRY(cudaMemcpy(&v, res.as<u8>() + 8, 4, cudaMemcpyDeviceToHost));
        *value = v;
        return SWC_OK;
    }
    switch (k) {
    case CRC32: st = crc32(d, n, res.as<u64>(), part.as<u64>(), 0); break;
    case BZIP2_CRC32: st = bzip2_crc32(d, n, res.as<u64>(), part.as<u64>(), 0); break;
    case CRC64: st = crc64(d, n, res.as<u64>(), part.as<u64>(), 0); break;
    default: st = adler32(d, n, res.as<u64>(), part.as<u64>(), 0); break;
    }
    if (st) return st;
    SWC_CUDA_TRY(cudaMemcpy(value, res.p, 8, cudaMemcpyDeviceToHost));
    return SWC_OK;
}

}  // namespace checks
}  // namespace swc

namespace {

inline uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }

// CRC-32 of the gzip *header* bytes for FHCRC (framing; <= a few hundred bytes) — GzipHeader.swift:191-198
uint32_t crc32_header(const uint8_t *p, size_t n) {
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; i++) {
        c ^= p[i];
        for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
    }
    return ~c;
}

// GzipHeader.init(_: LsbBitReader), GzipHeader.swift:68-199 — one pass over the member's header that both validates it and
// records where its fields are.  Written against the Swift source (guards in source order, one `Cursor` read per
// `reader.byte()`); `left()` is `reader.bytesLeft`.  Reads the reference performs WITHOUT a guard trap there (BitByteData
// precondition): they report SWC_ERR_REFERENCE_TRAP here.
struct Cursor {
    const uint8_t *base; size_t n, at;
    size_t left() const { return n - at; }
    bool finished() const { return at >= n; }
};

int gzip_parse(const uint8_t *in, size_t n, size_t member_off, swc_gzip_header *h) {
    if (member_off > n) return SWC_GZIP_WRONG_MAGIC;
    Cursor c{in, n, member_off};
    swc_gzip_header out;
    memset(&out, 0, sizeof(out));
    if (c.left() < 10) return SWC_GZIP_WRONG_MAGIC;                                  // :70-71
    if (in[c.at] != 0x1f || in[c.at + 1] != 0x8b) return SWC_GZIP_WRONG_MAGIC;       // :74-76 (uint16 == 0x8b1f)
    if (in[c.at + 2] != 8) return SWC_GZIP_WRONG_COMPRESSION_METHOD;                 // :80-82
    out.compression_method = 8;
    const unsigned flg = in[c.at + 3];
    if ((flg & 0xE0) != 0) return SWC_GZIP_WRONG_FLAGS;                              // :86-88
    out.modification_time = (uint32_t)in[c.at + 4] | (uint32_t)in[c.at + 5] << 8 | (uint32_t)in[c.at + 6] << 16 | (uint32_t)in[c.at + 7] << 24;
    out.os_type = in[c.at + 9];                                                      // :103-105 (XFL at +8 is only hashed)
    out.is_text_file = (flg & 0x01) != 0;                                            // :107
    c.at += 10;
    if (flg & 0x04) {                                                                // FEXTRA :111-155
        if (c.left() < 2) return SWC_GZIP_WRONG_MAGIC;                               // :112-113
        size_t xlen = (size_t)in[c.at] | (size_t)in[c.at + 1] << 8;
        c.at += 2;
        if (!(c.left() >= xlen && xlen >= 4)) return SWC_GZIP_WRONG_MAGIC;           // :123-124
        out.extra_off = c.at; out.extra_len = xlen;
        long remaining = (long)xlen;                                                 // the reference's signed `xlen` countdown
        while (remaining > 0) {                                                      // :125
            // si1, si2, two length bytes: four unguarded reader.byte() calls (:126-140)
            if (c.left() < 2) return SWC_ERR_REFERENCE_TRAP;
            if (in[c.at + 1] == 0) return SWC_GZIP_WRONG_FLAGS;                      // :131-132 (checked before the length is read)
            if (c.left() < 4) return SWC_ERR_REFERENCE_TRAP;
            const long len = (long)in[c.at + 2] | (long)in[c.at + 3] << 8;
            c.at += 4;
            remaining -= 4;                                                          // :141
            if (remaining < len) return SWC_GZIP_WRONG_MAGIC;                        // :145-146
            if ((long)c.left() < len) return SWC_ERR_REFERENCE_TRAP;                 // :148-152 unguarded payload reads
            c.at += (size_t)len;
            remaining -= len;                                                        // :154
        }
    }
    if (flg & 0x08) {                                                                // FNAME :159-172
        out.has_file_name = 1; out.file_name_off = c.at;
        for (;;) {
            if (c.finished()) return SWC_GZIP_WRONG_MAGIC;                           // :162-163
            if (in[c.at++] == 0) break;
        }
        out.file_name_len = c.at - 1 - out.file_name_off;
    }
    if (flg & 0x10) {                                                                // FCOMMENT :177-190
        out.has_comment = 1; out.comment_off = c.at;
        for (;;) {
            if (c.finished()) return SWC_GZIP_WRONG_MAGIC;                           // :180-181
            if (in[c.at++] == 0) break;
        }
        out.comment_len = c.at - 1 - out.comment_off;
    }
    if (flg & 0x02) {                                                                // FHCRC :194-200
        if (c.left() < 2) return SWC_GZIP_WRONG_MAGIC;
        const unsigned stored = (unsigned)in[c.at] | (unsigned)in[c.at + 1] << 8;
        if ((crc32_header(in + member_off, c.at - member_off) & 0xFFFFu) != stored) return SWC_GZIP_WRONG_HEADER_CRC;
        c.at += 2;
    }
    out.header_len = c.at - member_off;
    if (h) *h = out;
    return SWC_OK;
}

int gzip_header(const uint8_t *in, size_t n, size_t *off) {
    swc_gzip_header h;
    const int st = gzip_parse(in, n, *off, &h);
    if (st == SWC_OK) *off += h.header_len;
    return st;
}

// ZlibHeader.init(_: LsbBitReader), ZlibHeader.swift:47-93
int zlib_parse(const uint8_t *in, size_t n, swc_zlib_header *h) {
    if (n < 2) return SWC_ZLIB_WRONG_COMPRESSION_METHOD;                             // :49-50
    const unsigned cmf = in[0], flg = in[1];
    if ((cmf & 0xF) != 8) return SWC_ZLIB_WRONG_COMPRESSION_METHOD;                  // :56-58
    const unsigned cinfo = (cmf & 0xF0) >> 4;
    if (cinfo > 7) return SWC_ZLIB_WRONG_COMPRESSION_INFO;                           // :62-64
    // CompressionLevel(rawValue: (flags & 0xC0) >> 6) covers 0...3: it cannot fail (:80-82)
    if ((((unsigned)cmf << 8) + flg) % 31 != 0) return SWC_ZLIB_WRONG_FCHECK;        // :84-85 (Swift precedence: (cmf << 8) + flags)
    size_t len = 2;
    if ((flg & 0x20) != 0) {                                                         // FDICT: four bytes are skipped (:88-92)
        if (n - 2 < 4) return SWC_ZLIB_WRONG_FCHECK;
        len = 6;
    }
    if (h) { h->compression_method = 8; h->compression_level = (int32_t)((flg & 0xC0) >> 6); h->window_size = 1 << (cinfo + 8); h->header_len = len; }
    return SWC_OK;
}

// growable malloc'ed host buffer that is handed to the C caller without another copy (swc_free == free)
struct HostOut {
    uint8_t *p = nullptr;
    size_t len = 0, cap = 0;
    HostOut() {}
    HostOut(const HostOut &) = delete;
    HostOut &operator=(const HostOut &) = delete;
    ~HostOut() { free(p); }
    uint8_t *grow(size_t add) {                                     // nullptr when out of memory
        if (len + add > cap || !p) {
            size_t nc = cap * 2 > len + add ? cap * 2 : len + add;
            if (nc < 64) nc = 64;
            uint8_t *q = (uint8_t *)realloc(p, nc);
            if (!q) return nullptr;
            p = q; cap = nc;
        }
        uint8_t *r = p + len;
        len += add;
        return r;
    }
    int release_to(uint8_t **out, size_t *out_len) {
        if (!p && !grow(0)) return SWC_ERR_OUTPUT_OVERFLOW;
        *out = p; *out_len = len;
        p = nullptr; len = cap = 0;
        return SWC_OK;
    }
};

// processMember GzipArchive.swift:79-100; d_in holds the whole archive on the device
int gzip_member(const uint8_t *in, size_t n, const u8 *d_in, size_t *off, HostOut &out, bool *crc_error) {
    if (n - *off < 20) return SWC_GZIP_WRONG_MAGIC;
    int st = gzip_header(in, n, off);
    if (st) return st;
    UnitResult r;
    if ((st = deflate_unit_device(d_in, n, (size_t)*off * 8, r))) return st;
    if (r.status != SWC_OK) return r.status;
    size_t p = *off + (r.consumed + 7) / 8;                          // align()
    if (n - p < 8) return SWC_GZIP_WRONG_MAGIC;
    const uint32_t crc = rd32(in + p), isize = rd32(in + p + 4);
    p += 8;
    if ((uint32_t)r.out_len != isize) return SWC_GZIP_WRONG_ISIZE;
    u64 got = 0;
    if ((st = checks::check_device(checks::CRC32, r.out.as<u8>(), r.out_len, &got))) return st;
    *crc_error = (uint32_t)got != crc;
    uint8_t *dst = out.grow(r.out_len);
    if (!dst) return SWC_ERR_OUTPUT_OVERFLOW;
    if (r.out_len) SWC_CUDA_TRY(cudaMemcpy(dst, r.out.p, r.out_len, cudaMemcpyDeviceToHost));
    *off = p;
    return SWC_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Batch path for multi-member archives (BGZF and friends).  Members are independent Deflate streams, so candidate member
// starts are found by signature, every candidate is decoded speculatively in ONE batched launch (its input ends 8 bytes
// before the next candidate; its capacity is the ISIZE found there), and the in-order walk of GzipArchive.multiUnarchive
// (GzipArchive.swift:52-77) then only *validates*: member k is accepted iff it decoded cleanly, stopped exactly at that
// trailer and produced ISIZE bytes — exactly what the sequential decoder would have computed from the same start.  A
// member that does not validate (a signature look-alike inside its payload, ISIZE wrap-around, damage) is decoded by the
// sequential path, and the walk resynchronises on the candidate list afterwards.
static size_t inflate_scratch_bytes(u64 n, u64 out_total) { return swc_deflate_batch_scratch_bytes(n, out_total); }

struct Cand { size_t at, dstart, next; uint32_t crc, isize; long unit; };

int gzip_multi_batch(const uint8_t *in, size_t n, const u8 *d_in, HostOut &o, std::vector<size_t> &ends, std::vector<size_t> &starts, size_t *off, int *result, bool *stop) {
    const bool trace = getenv("SWC_TRACE") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t0 = now();
    std::vector<size_t> pos;
    {
        int st = checks::find_gzip_members(d_in, n, pos);              // device-side signature scan, sorted positions
        if (st) return st;
    }
    if (trace) fprintf(stderr, "[swc] gzip scan: %zu candidates, %.1f ms\n", pos.size(), now() - t0);
    if (pos.size() < 4 || pos[0] != 0) return SWC_OK;              // nothing to gain: the caller's loop handles it
    size_t free_b = 0, total_b = 0;
    SWC_CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    // a round needs its speculative outputs, the 4/3 x record scratch of the Deflate batch and the gather copy: ~3.3 x out_total
    const size_t budget = free_b / 5;
    cudaStream_t stream = 0;
    size_t idx = 0;
    while (idx < pos.size() && pos[idx] == *off && !*stop) {
        // ---- one round: candidates idx.. while the speculative outputs fit the budget
        std::vector<Cand> cs;
        std::vector<u64> h_inoff, h_inlen, h_outoff, h_outcap;
        size_t out_total = 0;
        for (size_t k = idx; k < pos.size() && cs.size() < (1u << 22); k++) {
            Cand c;
            c.at = pos[k]; c.next = k + 1 < pos.size() ? pos[k + 1] : n; c.unit = -1; c.crc = c.isize = 0;
            size_t d = c.at;
            const bool hdr_ok = gzip_header(in, c.next, &d) == SWC_OK;   // a real header never reaches the next member
            c.dstart = d;
            if (hdr_ok && c.next >= d + 8) {
                c.crc = rd32(in + c.next - 8); c.isize = rd32(in + c.next - 4);
                const size_t span = c.next - 8 - d;
                if ((u64)c.isize <= (u64)span * 1032 + 64) {            // Deflate cannot expand further than ~1032:1
                    const size_t cap16 = round16((size_t)c.isize);
                    if (out_total + cap16 > budget && !cs.empty()) break;
                    c.unit = (long)h_inoff.size();
                    h_inoff.push_back(d); h_inlen.push_back(span);
                    h_outoff.push_back(out_total); h_outcap.push_back(c.isize);
                    out_total += cap16;
                }
            }
            cs.push_back(c);
        }
        const size_t nu = h_inoff.size();
        std::vector<u64> h_outlen(nu), h_cons(nu);
        std::vector<int32_t> h_status(nu);
        std::vector<uint32_t> h_crc(nu);
        DevBuf meta, d_out;
        u64 *m = nullptr;
        if (nu) {
            // an allocation failure is not an error of the archive: hand the rest to the caller's sequential walk
            int st = meta.alloc(nu * 64);                               // 7 u64 arrays + status + crc
            if (st) return SWC_OK;
            if ((st = d_out.alloc(out_total + 64))) return SWC_OK;
            { void *scr = nullptr; if (scratch_get(inflate_scratch_bytes(nu, out_total), &scr, stream)) return SWC_OK; }
            m = meta.as<u64>();
            SWC_CUDA_TRY(cudaMemcpyAsync(m, h_inoff.data(), nu * 8, cudaMemcpyHostToDevice, stream));
            SWC_CUDA_TRY(cudaMemcpyAsync(m + nu, h_inlen.data(), nu * 8, cudaMemcpyHostToDevice, stream));
            SWC_CUDA_TRY(cudaMemcpyAsync(m + 2 * nu, h_outoff.data(), nu * 8, cudaMemcpyHostToDevice, stream));
            SWC_CUDA_TRY(cudaMemcpyAsync(m + 3 * nu, h_outcap.data(), nu * 8, cudaMemcpyHostToDevice, stream));
            int32_t *d_status = (int32_t *)(m + 7 * nu);
            uint32_t *d_crc = (uint32_t *)(d_status + nu);
            if ((st = deflate_batch_impl(d_in, m, m + nu, nullptr, d_out.as<u8>(), m + 2 * nu, m + 3 * nu, out_total,
                                         m + 4 * nu, m + 5 * nu, d_status, nu, nullptr, 0, stream))) return st;
            if ((st = checks::crc32_units(d_out.as<u8>(), m + 2 * nu, m + 4 * nu, d_status, d_crc, nu, stream))) return st;
            SWC_CUDA_TRY(cudaMemcpyAsync(h_outlen.data(), m + 4 * nu, nu * 8, cudaMemcpyDeviceToHost, stream));
            SWC_CUDA_TRY(cudaMemcpyAsync(h_cons.data(), m + 5 * nu, nu * 8, cudaMemcpyDeviceToHost, stream));
            SWC_CUDA_TRY(cudaMemcpyAsync(h_status.data(), d_status, nu * 4, cudaMemcpyDeviceToHost, stream));
            SWC_CUDA_TRY(cudaMemcpyAsync(h_crc.data(), d_crc, nu * 4, cudaMemcpyDeviceToHost, stream));
            SWC_CUDA_TRY(cudaStreamSynchronize(stream));
        }
        if (trace) fprintf(stderr, "[swc] gzip round: %zu units, %.1f MB out, decode+crc done at %.1f ms\n", nu, out_total / 1e6, now() - t0);
        // ---- the in-order walk: accept the longest validated prefix
        std::vector<u64> g_src, g_len, g_dst;
        size_t acc_bytes = 0, k = 0;
        for (; k < cs.size(); k++) {
            const Cand &c = cs[k];
            if (c.unit < 0) break;
            const size_t u = (size_t)c.unit;
            if (h_status[u] != SWC_OK || c.dstart + (h_cons[u] + 7) / 8 + 8 != c.next || h_outlen[u] != c.isize) break;
            g_src.push_back(h_outoff[u]); g_len.push_back(h_outlen[u]); g_dst.push_back(acc_bytes);
            acc_bytes += h_outlen[u];
            if (h_crc[u] != c.crc) { *result = SWC_GZIP_WRONG_CRC; *stop = true; k++; break; }   // the member is still returned
        }
        const size_t accepted = g_src.size();
        if (accepted) {
            DevBuf gm, d_g;
            int st = gm.alloc(accepted * 24);
            if (st) return SWC_OK;
            if ((st = d_g.alloc(acc_bytes + 16))) return SWC_OK;
            u64 *g = gm.as<u64>();
            SWC_CUDA_TRY(cudaMemcpyAsync(g, g_src.data(), accepted * 8, cudaMemcpyHostToDevice, stream));
            SWC_CUDA_TRY(cudaMemcpyAsync(g + accepted, g_len.data(), accepted * 8, cudaMemcpyHostToDevice, stream));
            SWC_CUDA_TRY(cudaMemcpyAsync(g + 2 * accepted, g_dst.data(), accepted * 8, cudaMemcpyHostToDevice, stream));
            if ((st = checks::gather_units(d_out.as<u8>(), g, g + accepted, d_g.as<u8>(), g + 2 * accepted, accepted, stream))) return st;
            const size_t base = o.len;
            uint8_t *dst = o.grow(acc_bytes);
            if (!dst) return SWC_ERR_OUTPUT_OVERFLOW;
            SWC_CUDA_TRY(cudaStreamSynchronize(stream));
            if ((st = copy_pageable(dst, d_g.p, acc_bytes, false))) return st;
            for (size_t i = 0; i < accepted; i++) { ends.push_back(base + g_dst[i] + g_len[i]); starts.push_back(cs[i].at); }
            *off = cs[accepted - 1].next;
            if (trace) fprintf(stderr, "[swc] gzip round: %zu accepted, gathered + copied back at %.1f ms\n", accepted, now() - t0);
        }
        if (*stop) return SWC_OK;
        idx += accepted;
        if (accepted == cs.size()) continue;                             // next round (or done)
        // ---- member idx did not validate: decode it the sequential way, then resynchronise on the candidate list
        bool crc_error = false;
        const size_t member_at = *off;
        int st = gzip_member(in, n, d_in, off, o, &crc_error);
        if (st) return st;
        ends.push_back(o.len); starts.push_back(member_at);
        if (crc_error) { *result = SWC_GZIP_WRONG_CRC; *stop = true; return SWC_OK; }
        idx = (size_t)(std::lower_bound(pos.begin(), pos.end(), *off) - pos.begin());
    }
    return SWC_OK;
}

int upload(DevBuf &d, const uint8_t *in, size_t n) {
    int st = d.alloc(round16(n) + 32);
    if (st) return st;
    return copy_pageable(d.p, in, n, true);
}

template <typename F>
int host_check(const uint8_t *in, size_t n, F fn) {
    if (ensure_device()) return SWC_ERR_NO_DEVICE;
    ApiLock api_lock;
    DevBuf d;
    int st = upload(d, in, n);
    if (st) return st;
    return fn(d.as<u8>());
}

}  // namespace

extern "C" {

int32_t swc_gzip_unarchive(const uint8_t *in, size_t in_len, uint8_t **out, size_t *out_len, size_t *consumed_bytes) {
    if (!out || !out_len) return SWC_ERR_INVALID_ARG;
    *out = nullptr; *out_len = 0;
    if (consumed_bytes) *consumed_bytes = 0;
    if (ensure_device()) return SWC_ERR_NO_DEVICE;
    ApiLock api_lock;
    DevBuf d_in;
    int st = upload(d_in, in, in_len);
    if (st) return st;
    HostOut o;
    size_t off = 0; bool crc_error = false;
    if ((st = gzip_member(in, in_len, d_in.as<u8>(), &off, o, &crc_error))) return st;
    if (consumed_bytes) *consumed_bytes = off;
    if ((st = o.release_to(out, out_len))) return st;
    return crc_error ? SWC_GZIP_WRONG_CRC : SWC_OK;
}

// GzipArchive.multiUnarchive GzipArchive.swift:52-77
static int gzip_multi(const uint8_t *in, size_t in_len, uint8_t **out, size_t *out_len, size_t **member_ends, size_t **member_in_off,
                      size_t *n_members) {
    if (!out || !out_len || !member_ends || !n_members) return SWC_ERR_INVALID_ARG;
    *out = nullptr; *out_len = 0; *member_ends = nullptr; *n_members = 0;
    if (member_in_off) *member_in_off = nullptr;
    if (ensure_device()) return SWC_ERR_NO_DEVICE;
    ApiLock api_lock;
    DevBuf d_in;
    int st = upload(d_in, in, in_len);
    if (st) return st;
    HostOut o;
    std::vector<size_t> ends, starts;
    size_t off = 0;
    int result = SWC_OK;
    bool stop = false;
    if ((st = gzip_multi_batch(in, in_len, d_in.as<u8>(), o, ends, starts, &off, &result, &stop))) return st;
    while (off < in_len && !stop) {
        bool crc_error = false;
        const size_t member_at = off;
        if ((st = gzip_member(in, in_len, d_in.as<u8>(), &off, o, &crc_error))) return st;
        ends.push_back(o.len); starts.push_back(member_at);
        if (crc_error) { result = SWC_GZIP_WRONG_CRC; break; }
    }
    if ((st = o.release_to(out, out_len))) return st;
    *member_ends = (size_t *)swc_alloc(sizeof(size_t) * (ends.size() + 1));
    for (size_t i = 0; i < ends.size(); i++) (*member_ends)[i] = ends[i];
    if (member_in_off) {
        *member_in_off = (size_t *)swc_alloc(sizeof(size_t) * (starts.size() + 1));
        for (size_t i = 0; i < starts.size(); i++) (*member_in_off)[i] = starts[i];
        (*member_in_off)[starts.size()] = off;
    }
    *n_members = ends.size();
    return result;
}

int32_t swc_gzip_multi_unarchive(const uint8_t *in, size_t in_len,
                                 uint8_t **out, size_t *out_len, size_t **member_ends, size_t *n_members) {
    return gzip_multi(in, in_len, out, out_len, member_ends, nullptr, n_members);
}

int32_t swc_gzip_multi_unarchive_members(const uint8_t *in, size_t in_len,
                                         uint8_t **out, size_t *out_len, size_t **member_ends, size_t **member_in_off,
                                         size_t *n_members) {
    if (!member_in_off) return SWC_ERR_INVALID_ARG;
    return gzip_multi(in, in_len, out, out_len, member_ends, member_in_off, n_members);
}

int32_t swc_gzip_header_parse(const uint8_t *in, size_t in_len, size_t member_off, swc_gzip_header *hdr) {
    if (!in || !hdr) return SWC_ERR_INVALID_ARG;
    return gzip_parse(in, in_len, member_off, hdr);
}

int32_t swc_zlib_header_parse(const uint8_t *in, size_t in_len, swc_zlib_header *hdr) {
    if (!in || !hdr) return SWC_ERR_INVALID_ARG;
    return zlib_parse(in, in_len, hdr);
}

int32_t swc_zlib_unarchive(const uint8_t *in, size_t n, uint8_t **out, size_t *out_len) {
    if (!out || !out_len) return SWC_ERR_INVALID_ARG;
    *out = nullptr; *out_len = 0;
    swc_zlib_header zh;
    { const int hst = zlib_parse(in, n, &zh); if (hst) return hst; }
    const size_t off = zh.header_len;
    if (ensure_device()) return SWC_ERR_NO_DEVICE;
    ApiLock api_lock;
    DevBuf d_in;
    int st = upload(d_in, in, n);
    if (st) return st;
    UnitResult r;
    if ((st = deflate_unit_device(d_in.as<u8>(), n, off * 8, r))) return st;
    if (r.status != SWC_OK) return r.status;
    const size_t p = off + (r.consumed + 7) / 8;
    int result = SWC_OK;
    if (n - p < 4) {
        result = SWC_ZLIB_WRONG_ADLER32;                             // ZlibArchive.swift:34-35 (payload still returned)
    } else {
        const uint32_t adler = (uint32_t)in[p] << 24 | (uint32_t)in[p + 1] << 16 | (uint32_t)in[p + 2] << 8 | in[p + 3];
        u64 got = 0;
        if ((st = checks::check_device(checks::ADLER32, r.out.as<u8>(), r.out_len, &got))) return st;
        if ((uint32_t)got != adler) result = SWC_ZLIB_WRONG_ADLER32;
    }
    if ((st = to_host_alloc(r.out.p, r.out_len, out, out_len))) return st;
    return result;
}

int32_t swc_crc32_batch(const uint8_t *in_base, const uint64_t *off, const uint64_t *len, const int32_t *status,
                        uint32_t *result, uint64_t n, void *cuda_stream) {
    if (n == 0) return SWC_OK;
    if (!in_base || !off || !len || !result) return SWC_ERR_INVALID_ARG;
    if (ensure_device()) return SWC_ERR_NO_DEVICE;
    return checks::crc32_units(in_base, off, len, status, result, n, (cudaStream_t)cuda_stream);
}
int32_t swc_xxh32_batch(const uint8_t *in_base, const uint64_t *off, const uint64_t *len,
                        uint32_t *result, uint64_t n, void *cuda_stream) {
    if (n == 0) return SWC_OK;
    if (!in_base || !off || !len || !result) return SWC_ERR_INVALID_ARG;
    if (ensure_device()) return SWC_ERR_NO_DEVICE;
    return checks::xxh32_batch(in_base, off, len, result, n, (cudaStream_t)cuda_stream);
}

int32_t swc_crc32(const uint8_t *in, size_t n, uint32_t *result) {
    return host_check(in, n, [&](const u8 *d) { u64 v = 0; int st = checks::check_device(checks::CRC32, d, n, &v); *result = (uint32_t)v; return st; });
}
int32_t swc_bzip2_crc32(const uint8_t *in, size_t n, uint32_t *result) {
    return host_check(in, n, [&](const u8 *d) { u64 v = 0; int st = checks::check_device(checks::BZIP2_CRC32, d, n, &v); *result = (uint32_t)v; return st; });
}
int32_t swc_crc64(const uint8_t *in, size_t n, uint64_t *result) {
    return host_check(in, n, [&](const u8 *d) { u64 v = 0; int st = checks::check_device(checks::CRC64, d, n, &v); *result = v; return st; });
}
int32_t swc_adler32(const uint8_t *in, size_t n, uint32_t *result) {
    return host_check(in, n, [&](const u8 *d) { u64 v = 0; int st = checks::check_device(checks::ADLER32, d, n, &v); *result = (uint32_t)v; return st; });
}
int32_t swc_xxh32(const uint8_t *in, size_t n, uint32_t *result) {
    return host_check(in, n, [&](const u8 *d) { u64 v = 0; int st = checks::check_device(checks::XXH32, d, n, &v); *result = (uint32_t)v; return st; });
}
int32_t swc_sha256(const uint8_t *in, size_t n, uint8_t digest[32]) {
    return host_check(in, n, [&](const u8 *d) {
        DevBuf dg;
        int st = dg.alloc(32);
        if (st) return st;
        if ((st = checks::sha256(d, n, dg.as<u8>(), 0))) return st;
        SWC_CUDA_TRY(cudaMemcpy(digest, dg.p, 32, cudaMemcpyDeviceToHost));
        return (int)SWC_OK;
    });
}

}  // extern "C"
