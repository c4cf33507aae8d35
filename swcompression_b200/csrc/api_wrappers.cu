// api_wrappers.cu — GzipArchive / ZlibArchive framing around the device Deflate decoder, and the checksum C ABI.
// Reference: Sources/GZip/GzipArchive.swift:38-100, GzipHeader.swift:68-199, Sources/Zlib/ZlibArchive.swift:25-42,
// ZlibHeader.swift:47-93.  Headers/trailers (tens of bytes) are parsed on the host; Deflate and CRC-32 / Adler-32 over
// the payload run on the device.
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

// GzipHeader.init(_:) GzipHeader.swift:68-199
int gzip_header(const uint8_t *in, size_t n, size_t *off) {
    size_t p = *off;
    if (n - p < 10) return SWC_GZIP_WRONG_MAGIC;
    if (in[p] != 0x1f || in[p + 1] != 0x8b) return SWC_GZIP_WRONG_MAGIC;
    if (in[p + 2] != 8) return SWC_GZIP_WRONG_COMPRESSION_METHOD;
    const unsigned flags = in[p + 3];
    if (flags & 0xE0) return SWC_GZIP_WRONG_FLAGS;
    const size_t hstart = p;
    p += 10;
    if (flags & 0x04) {                                             // FEXTRA
        if (n - p < 2) return SWC_GZIP_WRONG_MAGIC;
        long xlen = in[p] | in[p + 1] << 8; p += 2;
        if (!((long)(n - p) >= xlen && xlen >= 4)) return SWC_GZIP_WRONG_MAGIC;
        while (xlen > 0) {
            if (n - p < 4) return SWC_ERR_REFERENCE_TRAP;           // unguarded reads past the end trap in the reference
            if (in[p + 1] == 0) return SWC_GZIP_WRONG_FLAGS;
            long len = in[p + 2] | in[p + 3] << 8; p += 4;
            xlen -= 4;
            if (xlen < len) return SWC_GZIP_WRONG_MAGIC;
            if ((long)(n - p) < len) return SWC_ERR_REFERENCE_TRAP;
            p += (size_t)len; xlen -= len;
        }
    }
    for (unsigned bit : {0x08u, 0x10u}) {                           // FNAME, FCOMMENT
        if (!(flags & bit)) continue;
        for (;;) {
            if (p >= n) return SWC_GZIP_WRONG_MAGIC;
            if (in[p++] == 0) break;
        }
    }
    if (flags & 0x02) {                                             // FHCRC
        if (n - p < 2) return SWC_GZIP_WRONG_MAGIC;
        unsigned crc16 = in[p] | in[p + 1] << 8;
        if ((crc32_header(in + hstart, p - hstart) & 0xFFFF) != crc16) return SWC_GZIP_WRONG_HEADER_CRC;
        p += 2;
    }
    *off = p;
    return SWC_OK;
}

// processMember GzipArchive.swift:79-100; d_in holds the whole archive on the device
int gzip_member(const uint8_t *in, size_t n, const u8 *d_in, size_t *off, std::vector<uint8_t> &out, bool *crc_error) {
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
    size_t base = out.size();
    out.resize(base + r.out_len);
    if (r.out_len) SWC_CUDA_TRY(cudaMemcpy(out.data() + base, r.out.p, r.out_len, cudaMemcpyDeviceToHost));
    *off = p;
    return SWC_OK;
}

int give(const std::vector<uint8_t> &v, uint8_t **out, size_t *out_len) {
    uint8_t *h = (uint8_t *)swc_alloc(v.size());
    if (!h) return SWC_ERR_OUTPUT_OVERFLOW;
    if (!v.empty()) memcpy(h, v.data(), v.size());
    *out = h; *out_len = v.size();
    return SWC_OK;
}

int upload(DevBuf &d, const uint8_t *in, size_t n) {
    int st = d.alloc(round16(n) + 32);
    if (st) return st;
    if (n) SWC_CUDA_TRY(cudaMemcpy(d.p, in, n, cudaMemcpyHostToDevice));
    return SWC_OK;
}

template <typename F>
int host_check(const uint8_t *in, size_t n, F fn) {
    if (ensure_device()) return SWC_ERR_NO_DEVICE;
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
    DevBuf d_in;
    int st = upload(d_in, in, in_len);
    if (st) return st;
    std::vector<uint8_t> o;
    size_t off = 0; bool crc_error = false;
    if ((st = gzip_member(in, in_len, d_in.as<u8>(), &off, o, &crc_error))) return st;
    if (consumed_bytes) *consumed_bytes = off;
    if ((st = give(o, out, out_len))) return st;
    return crc_error ? SWC_GZIP_WRONG_CRC : SWC_OK;
}

int32_t swc_gzip_multi_unarchive(const uint8_t *in, size_t in_len,
                                 uint8_t **out, size_t *out_len, size_t **member_ends, size_t *n_members) {
    if (!out || !out_len || !member_ends || !n_members) return SWC_ERR_INVALID_ARG;
    *out = nullptr; *out_len = 0; *member_ends = nullptr; *n_members = 0;
    if (ensure_device()) return SWC_ERR_NO_DEVICE;
    DevBuf d_in;
    int st = upload(d_in, in, in_len);
    if (st) return st;
    std::vector<uint8_t> o;
    std::vector<size_t> ends;
    size_t off = 0;
    int result = SWC_OK;
    while (off < in_len) {
        bool crc_error = false;
        if ((st = gzip_member(in, in_len, d_in.as<u8>(), &off, o, &crc_error))) return st;
        ends.push_back(o.size());
        if (crc_error) { result = SWC_GZIP_WRONG_CRC; break; }
    }
    if ((st = give(o, out, out_len))) return st;
    *member_ends = (size_t *)swc_alloc(sizeof(size_t) * (ends.size() + 1));
    for (size_t i = 0; i < ends.size(); i++) (*member_ends)[i] = ends[i];
    *n_members = ends.size();
    return result;
}

int32_t swc_zlib_unarchive(const uint8_t *in, size_t n, uint8_t **out, size_t *out_len) {
    if (!out || !out_len) return SWC_ERR_INVALID_ARG;
    *out = nullptr; *out_len = 0;
    if (n < 2) return SWC_ZLIB_WRONG_COMPRESSION_METHOD;             // ZlibHeader.swift:49
    const unsigned cmf = in[0], flags = in[1];
    if ((cmf & 0xF) != 8) return SWC_ZLIB_WRONG_COMPRESSION_METHOD;
    if (((cmf & 0xF0) >> 4) > 7) return SWC_ZLIB_WRONG_COMPRESSION_INFO;
    if (((cmf << 8) + flags) % 31 != 0) return SWC_ZLIB_WRONG_FCHECK;
    size_t off = 2;
    if ((flags & 0x20) >> 5) { if (n - off < 4) return SWC_ZLIB_WRONG_FCHECK; off += 4; }
    if (ensure_device()) return SWC_ERR_NO_DEVICE;
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
