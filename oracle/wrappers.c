/* wrappers.c — ORACLE (test infrastructure): GZip / Zlib / XZ framing around the codecs.
 *   Sources/GZip/GzipArchive.swift:38-100, GzipHeader.swift:68-199
 *   Sources/Zlib/ZlibArchive.swift:25-42, ZlibHeader.swift:47-93
 *   Sources/XZ/XZArchive.swift:27-218, XZBlock.swift:18-97, XZStreamHeader.swift:33-57,
 *   Sources/XZ/LittleEndianByteReader+XZ.swift:10-30, Sources/Common/DeltaFilter.swift:11-32
 * Unguarded reads past the end of input are BitByteData precondition traps in the reference -> SWC_ERR_REFERENCE_TRAP. */
#include "swco.h"

/* ------------------------------------------------------------------ GZip */
/* GzipHeader.init(_:) GzipHeader.swift:68-199. *off is a byte offset into in[0..n). */
static int gzip_header(const uint8_t *in, size_t n, size_t *off) {
    size_t p = *off;
    if (n - p < 10) return SWC_GZIP_WRONG_MAGIC;
    if (in[p] != 0x1f || in[p + 1] != 0x8b) return SWC_GZIP_WRONG_MAGIC;
    if (in[p + 2] != 8) return SWC_GZIP_WRONG_COMPRESSION_METHOD;
    unsigned flags = in[p + 3];
    if (flags & 0xE0) return SWC_GZIP_WRONG_FLAGS;
    const size_t hstart = p;
    p += 10;
    if (flags & 0x04) {                                            /* FEXTRA :111-157 */
        if (n - p < 2) return SWC_GZIP_WRONG_MAGIC;
        long xlen = in[p] | in[p + 1] << 8; p += 2;
        if (!((long)(n - p) >= xlen && xlen >= 4)) return SWC_GZIP_WRONG_MAGIC;
        while (xlen > 0) {
            /* the reference reads si1, si2, len (4 bytes) without re-checking xlen >= 4; input length was checked
               against the initial xlen only, so a short tail can run past: trap */
            if (n - p < 4) return SWC_ERR_REFERENCE_TRAP;
            unsigned si2 = in[p + 1];
            if (si2 == 0) return SWC_GZIP_WRONG_FLAGS;
            long len = in[p + 2] | in[p + 3] << 8; p += 4;
            xlen -= 4;
            if (xlen < len) return SWC_GZIP_WRONG_MAGIC;
            if ((long)(n - p) < len) return SWC_ERR_REFERENCE_TRAP;
            p += (size_t)len;
            xlen -= len;
        }
    }
    for (int pass = 0; pass < 2; pass++) {                         /* FNAME :159-172, FCOMMENT :175-188 */
        if (!(flags & (pass == 0 ? 0x08 : 0x10))) continue;
        for (;;) {
            if (p >= n) return SWC_GZIP_WRONG_MAGIC;
            if (in[p++] == 0) break;
        }
    }
    if (flags & 0x02) {                                            /* FHCRC :191-198 */
        if (n - p < 2) return SWC_GZIP_WRONG_MAGIC;
        unsigned crc16 = in[p] | in[p + 1] << 8;
        if ((swco_crc32(in + hstart, p - hstart, 0) & 0xFFFF) != crc16) return SWC_GZIP_WRONG_HEADER_CRC;
        p += 2;
    }
    *off = p;
    return SWC_OK;
}

/* processMember GzipArchive.swift:79-100. crc_error is reported separately (checked by the caller after the member
 * has been appended, :44,:72). */
static int gzip_member(const uint8_t *in, size_t n, size_t *off, swco_buf *out, int *crc_error) {
    if (n - *off < 20) return SWC_GZIP_WRONG_MAGIC;                /* reader is byte aligned by construction */
    int st = gzip_header(in, n, off);
    if (st) return st;
    size_t mstart = out->len;
    uint64_t used_bits = 0;
    st = swco_deflate_decompress(in, n, (uint64_t)*off * 8, out, &used_bits);
    if (st) return st;
    size_t p = *off + (size_t)((used_bits + 7) / 8);               /* align() */
    if (n - p < 8) return SWC_GZIP_WRONG_MAGIC;
    uint32_t crc = (uint32_t)in[p] | (uint32_t)in[p + 1] << 8 | (uint32_t)in[p + 2] << 16 | (uint32_t)in[p + 3] << 24;
    uint32_t isize = (uint32_t)in[p + 4] | (uint32_t)in[p + 5] << 8 | (uint32_t)in[p + 6] << 16 | (uint32_t)in[p + 7] << 24;
    p += 8;
    if ((uint32_t)(out->len - mstart) != isize) return SWC_GZIP_WRONG_ISIZE;
    *crc_error = swco_crc32(out->data + mstart, out->len - mstart, 0) != crc;
    *off = p;
    return SWC_OK;
}

int swco_gzip_unarchive(const uint8_t *in, size_t n, swco_buf *out, size_t *consumed) {
    size_t off = 0; int crc_error = 0;
    int st = gzip_member(in, n, &off, out, &crc_error);
    if (consumed) *consumed = off;
    if (st) return st;
    return crc_error ? SWC_GZIP_WRONG_CRC : SWC_OK;
}

int swco_gzip_multi_unarchive(const uint8_t *in, size_t n, swco_buf *out, size_t *ends, size_t max_n, size_t *cnt_out) {
    size_t off = 0, cnt = 0;
    if (cnt_out) *cnt_out = 0;
    while (off < n) {
        int crc_error = 0;
        int st = gzip_member(in, n, &off, out, &crc_error);
        if (st) return st;
        if (cnt < max_n) ends[cnt] = out->len;
        cnt++;
        if (cnt_out) *cnt_out = cnt;
        if (crc_error) return SWC_GZIP_WRONG_CRC;
    }
    return SWC_OK;
}

/* ------------------------------------------------------------------ Zlib */
int swco_zlib_unarchive(const uint8_t *in, size_t n, swco_buf *out) {
    if (n < 2) return SWC_ZLIB_WRONG_COMPRESSION_METHOD;           /* ZlibHeader.swift:49 */
    unsigned cmf = in[0], flags = in[1];
    if ((cmf & 0xF) != 8) return SWC_ZLIB_WRONG_COMPRESSION_METHOD;
    if (((cmf & 0xF0) >> 4) > 7) return SWC_ZLIB_WRONG_COMPRESSION_INFO;
    /* CompressionLevel(rawValue: 0...3) always succeeds: wrongCompressionLevel is unreachable (:79) */
    if (((cmf << 8) + flags) % 31 != 0) return SWC_ZLIB_WRONG_FCHECK;  /* Swift precedence: (cmf << 8) + flags */
    size_t off = 2;
    if ((flags & 0x20) >> 5) { if (n - off < 4) return SWC_ZLIB_WRONG_FCHECK; off += 4; }
    uint64_t used_bits = 0;
    size_t start = out->len;
    int st = swco_deflate_decompress(in, n, (uint64_t)off * 8, out, &used_bits);
    if (st) return st;
    size_t p = off + (size_t)((used_bits + 7) / 8);
    if (n - p < 4) return SWC_ZLIB_WRONG_ADLER32;                  /* ZlibArchive.swift:34 */
    uint32_t adler = (uint32_t)in[p] << 24 | (uint32_t)in[p + 1] << 16 | (uint32_t)in[p + 2] << 8 | in[p + 3];
    if (swco_adler32(out->data + start, out->len - start) != adler) return SWC_ZLIB_WRONG_ADLER32;
    return SWC_OK;
}

/* ------------------------------------------------------------------ Delta filter */
int swco_delta_decode(const uint8_t *in, size_t n, int distance, swco_buf *out) {   /* DeltaFilter.swift:11-32 */
    uint8_t delta[256] = {0};
    int pos = 0;
    for (size_t i = 0; i < n; i++) {
        uint8_t tmp = (uint8_t)(in[i] + delta[(distance + pos) % 256]);
        delta[pos] = tmp;
        if (swco_buf_push(out, tmp)) return SWC_ERR_OUTPUT_OVERFLOW;
        pos = pos == 0 ? 255 : pos - 1;
    }
    return SWC_OK;
}

/* ------------------------------------------------------------------ XZ */
typedef struct { const uint8_t *in; size_t n, off; } br_t;
#define NEED(r, k) do { if ((r)->n - (r)->off < (size_t)(k)) return SWC_ERR_REFERENCE_TRAP; } while (0)
static inline uint32_t le32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }

static int multibyte(br_t *r, int64_t *val) {                       /* LittleEndianByteReader+XZ.swift:10-30 */
    NEED(r, 1);
    int i = 1;
    int64_t result = r->in[r->off++];
    if (result <= 127) { *val = result; return SWC_OK; }
    result &= 0x7F;
    for (;;) {
        NEED(r, 1);
        unsigned b = r->in[r->off++];
        if (i >= 9 || b == 0) return SWC_XZ_MULTI_BYTE_INTEGER_ERROR;
        result += (int64_t)(b & 0x7F) << (7 * i);
        i++;
        if ((b & 0x80) == 0) break;
    }
    *val = result;
    return SWC_OK;
}

static int check_size(int type) { return type == 0 ? 0 : type == 1 ? 4 : type == 4 ? 8 : 32; }

/* XZBlock.init XZBlock.swift:18-97; the block's data is appended to `out`. */
static int xz_block(br_t *r, unsigned hsize_byte, int csize, swco_buf *out, int64_t *unpadded, int64_t *uncomp) {
    size_t hstart = r->off - 1;
    size_t real = ((size_t)hsize_byte + 1) * 4;
    NEED(r, 1);
    unsigned flags = r->in[r->off++];
    int nfilters = (flags & 0x03) + 1;
    if (flags & 0x3C) return SWC_XZ_WRONG_FIELD;
    int64_t comp_size = -1, uncomp_size = -1;
    int st;
    if (flags & 0x40) { if ((st = multibyte(r, &comp_size))) return st; }
    if (flags & 0x80) { if ((st = multibyte(r, &uncomp_size))) return st; }
    int kinds[4], params[4];
    for (int f = 0; f < nfilters; f++) {
        int64_t id, psz;
        if ((st = multibyte(r, &id))) return st;
        if ((uint64_t)id >= 0x4000000000000000ull) return SWC_XZ_WRONG_FILTER_ID;
        if (id == 0x21) {
            if ((st = multibyte(r, &psz))) return st;
            if (psz != 1) return SWC_LZMA2_WRONG_DICTIONARY_SIZE;
            NEED(r, 1);
            kinds[f] = 0x21; params[f] = r->in[r->off++];
        } else if (id == 0x03) {
            if ((st = multibyte(r, &psz))) return st;
            if (psz != 1) return SWC_XZ_WRONG_FIELD;
            NEED(r, 1);
            kinds[f] = 0x03; params[f] = ((r->in[r->off++] + 1) & 0xFF);
        } else {
            return SWC_XZ_WRONG_FILTER_ID;
        }
    }
    while ((int64_t)(r->off - hstart) < (int64_t)real - 4) {         /* header padding :65-69 */
        NEED(r, 1);
        if (r->in[r->off++] != 0) return SWC_XZ_WRONG_PADDING;
    }
    NEED(r, 4);
    uint32_t hcrc = le32(r->in + r->off);
    if (r->n - hstart < real - 4) return SWC_ERR_REFERENCE_TRAP;
    if (swco_crc32(r->in + hstart, real - 4, 0) != hcrc) return SWC_XZ_WRONG_INFO_CRC;
    r->off = hstart + (real - 4) + 4;                                /* :72-75 */

    size_t data_start = r->off;
    /* filters.reversed().reduce(byteReader): the LAST filter reads the archive reader, earlier ones read its output */
    swco_buf cur; swco_buf_init(&cur);
    int have_cur = 0;
    st = SWC_OK;
    for (int f = nfilters - 1; f >= 0 && st == SWC_OK; f--) {
        swco_buf next; swco_buf_init(&next);
        const uint8_t *src = have_cur ? cur.data : r->in + r->off;
        size_t src_n = have_cur ? cur.len : r->n - r->off;
        size_t used = 0;
        if (kinds[f] == 0x21) st = swco_lzma2_decompress_raw(src, src_n, (uint8_t)params[f], &next, &used);
        else { st = swco_delta_decode(src, src_n, params[f], &next); used = src_n; }
        if (!have_cur) r->off += used;
        swco_buf_free(&cur);
        cur = next; have_cur = 1;
    }
    if (st) { swco_buf_free(&cur); return st; }
    if (!((comp_size < 0 || comp_size == (int64_t)(r->off - data_start)) &&
          (uncomp_size < 0 || uncomp_size == (int64_t)cur.len))) { swco_buf_free(&cur); return SWC_XZ_WRONG_DATA_SIZE; }
    int64_t unp = (int64_t)(r->off - hstart);
    if (unp % 4 != 0) {
        int pad = 4 - (int)(unp % 4);
        for (int i = 0; i < pad; i++) {
            if (r->n - r->off < 1) { swco_buf_free(&cur); return SWC_ERR_REFERENCE_TRAP; }
            if (r->in[r->off++] != 0) { swco_buf_free(&cur); return SWC_XZ_WRONG_PADDING; }
        }
    }
    *uncomp = (int64_t)cur.len;
    *unpadded = unp + csize;
    int rc = swco_buf_append(out, cur.data, cur.len);
    swco_buf_free(&cur);
    return rc ? SWC_ERR_OUTPUT_OVERFLOW : SWC_OK;
}

/* processStream XZArchive.swift:90-130 (+ processIndex :132-167, processFooter :169-192) */
static int xz_stream(br_t *r, swco_buf *out, int *check_error) {
    static const uint8_t magic[6] = {0xFD, 0x37, 0x7A, 0x58, 0x5A, 0x00};
    NEED(r, 12);
    if (memcmp(r->in + r->off, magic, 6) != 0) return SWC_XZ_WRONG_MAGIC;
    const uint8_t *fl = r->in + r->off + 6;
    if (swco_crc32(fl, 2, 0) != le32(r->in + r->off + 8)) return SWC_XZ_WRONG_INFO_CRC;
    if (!(fl[0] == 0 && (fl[1] & 0xF0) == 0)) return SWC_XZ_WRONG_FIELD;
    int ctype = fl[1] & 0xF;
    if (!(ctype == 0 || ctype == 1 || ctype == 4 || ctype == 0x0A)) return SWC_XZ_WRONG_FIELD;
    r->off += 12;

    int64_t infos[2][4096]; /* bounded for the oracle; streams in tests have few blocks */
    int nblocks = 0;
    int64_t index_size = -1;
    *check_error = 0;
    for (;;) {
        NEED(r, 1);
        unsigned hs = r->in[r->off++];
        if (hs == 0) {                                               /* processIndex */
            size_t istart = r->off - 1;
            int64_t v; int st;
            if ((st = multibyte(r, &v))) return st;
            if (v != nblocks) return SWC_XZ_WRONG_FIELD;
            for (int b = 0; b < nblocks; b++) {
                if ((st = multibyte(r, &v))) return st;
                if (v != infos[0][b]) return SWC_XZ_WRONG_FIELD;
                if ((st = multibyte(r, &v))) return st;
                if (v != infos[1][b]) return SWC_XZ_WRONG_DATA_SIZE;
            }
            int64_t isz = (int64_t)(r->off - istart);
            if (isz % 4 != 0) {
                int pad = 4 - (int)(isz % 4);
                for (int i = 0; i < pad; i++) { NEED(r, 1); if (r->in[r->off++] != 0) return SWC_XZ_WRONG_PADDING; isz++; }
            }
            NEED(r, 4);
            uint32_t icrc = le32(r->in + r->off);
            if (swco_crc32(r->in + istart, (size_t)isz, 0) != icrc) return SWC_XZ_WRONG_INFO_CRC;
            r->off = istart + (size_t)isz + 4;
            index_size = isz + 4;
            break;
        }
        if (nblocks >= 4096) return SWC_ERR_UNSUPPORTED;
        size_t bstart = out->len;
        int64_t unp = 0, unc = 0;
        int st = xz_block(r, hs, check_size(ctype), out, &unp, &unc);
        if (st) return st;
        const uint8_t *bd = out->data + bstart; size_t bl = out->len - bstart;
        if (ctype == 1) {
            NEED(r, 4);
            uint32_t c = le32(r->in + r->off); r->off += 4;
            if (swco_crc32(bd, bl, 0) != c) { *check_error = 1; return SWC_OK; }
        } else if (ctype == 4) {
            NEED(r, 8);
            uint64_t c = (uint64_t)le32(r->in + r->off) | (uint64_t)le32(r->in + r->off + 4) << 32; r->off += 8;
            if (swco_crc64(bd, bl) != c) { *check_error = 1; return SWC_OK; }
        } else if (ctype == 0x0A) {
            NEED(r, 32);
            uint8_t dg[32]; swco_sha256(bd, bl, dg);
            int bad = memcmp(dg, r->in + r->off, 32) != 0; r->off += 32;
            if (bad) { *check_error = 1; return SWC_OK; }
        }
        infos[0][nblocks] = unp; infos[1][nblocks] = unc; nblocks++;
    }
    /* processFooter */
    NEED(r, 12);
    uint32_t fcrc = le32(r->in + r->off);
    int64_t backward = ((int64_t)le32(r->in + r->off + 4) + 1) * 4;
    unsigned fflags = r->in[r->off + 8] | r->in[r->off + 9] << 8;
    if (swco_crc32(r->in + r->off + 4, 6, 0) != fcrc) return SWC_XZ_WRONG_INFO_CRC;
    if (backward != index_size) return SWC_XZ_WRONG_FIELD;
    if (!((fflags & 0xFF) == 0 && ((fflags & 0xF00) >> 8) == (unsigned)ctype && (fflags & 0xF000) == 0)) return SWC_XZ_WRONG_FIELD;
    if (!(r->in[r->off + 10] == 0x59 && r->in[r->off + 11] == 0x5A)) return SWC_XZ_WRONG_MAGIC;
    r->off += 12;
    return SWC_OK;
}

static int xz_padding(br_t *r) {                                      /* processPadding :194-218 */
    if (r->off >= r->n) return SWC_OK;
    int padding = 0;
    for (;;) {
        unsigned b = r->in[r->off++];
        if (b != 0) { if (padding % 4 != 0) return SWC_XZ_WRONG_PADDING; break; }
        if (r->off >= r->n) { if (padding % 4 != 3) return SWC_XZ_WRONG_PADDING; return SWC_OK; }
        padding++;
    }
    r->off -= 1;
    return SWC_OK;
}

int swco_xz_split_unarchive(const uint8_t *in, size_t n, swco_buf *out, size_t *ends, size_t max_n, size_t *cnt_out) {
    br_t r = {in, n, 0};
    size_t cnt = 0;
    if (cnt_out) *cnt_out = 0;
    while (r.off < r.n) {
        if (r.n - r.off < 32) return SWC_XZ_WRONG_MAGIC;
        int check_error = 0;
        int st = xz_stream(&r, out, &check_error);
        if (st) return st;
        if (ends && cnt < max_n) ends[cnt] = out->len;
        cnt++;
        if (cnt_out) *cnt_out = cnt;
        if (check_error) return SWC_XZ_WRONG_CHECK;
        if ((st = xz_padding(&r))) return st;
    }
    return SWC_OK;
}

int swco_xz_unarchive(const uint8_t *in, size_t n, swco_buf *out) {
    return swco_xz_split_unarchive(in, n, out, NULL, 0, NULL);
}
