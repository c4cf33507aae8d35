/* lz4.c — ORACLE (test infrastructure): restatement of Sources/LZ4/LZ4.swift:73-413.
 * Line references are to that file. Errors are the reference's DataError cases (SWC_DATA_*). */
#include "swco.h"

static inline uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
static inline uint64_t rd64(const uint8_t *p) { return (uint64_t)rd32(p) | (uint64_t)rd32(p + 4) << 32; }
static inline int is_magic(uint32_t v) { return v == 0x184D2204u || v == 0x184C2102u || (v >= 0x184D2A50u && v <= 0x184D2A5Fu); }

/* process(block:_:) :332-413.  The reference prepends `dict` to a scratch array, decodes, strips it again (:334,:412);
 * the decoded block is appended to `out`. */
int swco_lz4_block(const uint8_t *in, size_t n, const uint8_t *dict, size_t dict_len, swco_buf *out) {
    swco_buf w; swco_buf_init(&w);
    int status = SWC_OK;
#define FAIL(c) do { status = (c); goto done; } while (0)
    if (dict_len && swco_buf_append(&w, dict, dict_len)) FAIL(SWC_ERR_OUTPUT_OVERFLOW);
    const size_t out_start = w.len;
    size_t ip = 0;
    long long sequence_count = 0;
    long long last_match_start = -1;
    for (;;) {
        sequence_count++;
        if (n - ip < 1) FAIL(SWC_DATA_TRUNCATED);                                 /* :343 */
        unsigned token = in[ip++];
        uint64_t lit = token >> 4;
        if (lit == 15) {                                                          /* :347-363 */
            for (;;) {
                if (n - ip < 1) FAIL(SWC_DATA_TRUNCATED);
                unsigned b = in[ip++];
                lit += b;                     /* Int overflow (unsupportedFeature) needs > 2^55 input bytes: unreachable */
                if (b != 255) break;
            }
        }
        if (n - ip < lit) FAIL(SWC_DATA_TRUNCATED);                               /* :364 */
        if (swco_buf_append(&w, in + ip, (size_t)lit)) FAIL(SWC_ERR_OUTPUT_OVERFLOW);
        ip += (size_t)lit;
        if (ip == n) {                                                            /* :369-377 reader.isFinished */
            if (!(lit >= 5 || sequence_count == 1)) FAIL(SWC_DATA_CORRUPTED);
            if (!((long long)w.len - last_match_start >= 12 || last_match_start == -1)) FAIL(SWC_DATA_CORRUPTED);
            break;
        }
        if (n - ip < 2) FAIL(SWC_DATA_TRUNCATED);                                 /* :379 */
        size_t offset = (size_t)in[ip] | (size_t)in[ip + 1] << 8;
        ip += 2;
        if (!(offset > 0 && offset <= w.len)) FAIL(SWC_DATA_CORRUPTED);           /* :383 */
        uint64_t mlen = 4 + (token & 0xF);
        if (mlen == 19) {                                                         /* :386-401 */
            for (;;) {
                if (n - ip < 1) FAIL(SWC_DATA_TRUNCATED);
                unsigned b = in[ip++];
                mlen += b;
                if (b != 255) break;
            }
        }
        last_match_start = (long long)w.len;                                      /* :405 */
        if (swco_buf_reserve(&w, (size_t)mlen)) FAIL(SWC_ERR_OUTPUT_OVERFLOW);
        size_t src = w.len - offset;
        for (uint64_t i = 0; i < mlen; i++) w.data[w.len + i] = w.data[src + i]; /* :407-409 */
        w.len += (size_t)mlen;
    }
    if (swco_buf_append(out, w.data + out_start, w.len - out_start)) status = SWC_ERR_OUTPUT_OVERFLOW;
done:
    swco_buf_free(&w);
    return status;
#undef FAIL
}

/* process(legacyFrame:) :160-186; `in` starts right after the magic. */
static int legacy_frame(const uint8_t *in, size_t n, swco_buf *out, size_t *used) {
    size_t off = 0;
    while (off < n) {
        if (n - off < 4) return SWC_DATA_TRUNCATED;
        uint32_t raw = rd32(in + off); off += 4;
        if (is_magic(raw)) { off -= 4; break; }                                   /* :168-171 */
        size_t bs = raw;
        if (n - off < bs) return SWC_DATA_TRUNCATED;
        int st = swco_lz4_block(in + off, bs, NULL, 0, out);
        if (st) return st;
        off += bs;
    }
    *used = off;
    return SWC_OK;
}

/* process(frame:_:_:) :188-330; `in` starts right after the magic. */
static int frame(const uint8_t *in, size_t n, const uint8_t *dict, size_t dict_len, int have_dict,
                 int has_ext_id, uint32_t ext_id, swco_buf *out, size_t *used) {
    if (n < 7) return SWC_DATA_TRUNCATED;                                         /* :191 */
    size_t off = 0;
    unsigned flg = in[off++];
    if (!(((flg & 0xC0) >> 6) == 1 && (flg & 0x2) == 0)) return SWC_DATA_CORRUPTED;
    int independent = (flg & 0x20) != 0, block_ck = (flg & 0x10) != 0, csize_p = (flg & 0x8) != 0;
    int cck = (flg & 0x4) != 0, dictid_p = (flg & 1) != 0;
    unsigned bd = in[off++];
    size_t max_block;
    switch (bd) {                                                                 /* :214-228 */
    case 0x40: max_block = 64u << 10; break;
    case 0x50: max_block = 256u << 10; break;
    case 0x60: max_block = 1u << 20; break;
    case 0x70: max_block = 4u << 20; break;
    default: return SWC_DATA_CORRUPTED;
    }
    uint64_t content_size = 0;
    if (csize_p) {
        if (n - off < 13) return SWC_DATA_TRUNCATED;
        content_size = rd64(in + off); off += 8;
        if (content_size > (uint64_t)INT64_MAX) return SWC_DATA_UNSUPPORTED_FEATURE;
    }
    if (dictid_p) {                                                               /* :247-270 */
        if (!have_dict) return SWC_DATA_CORRUPTED;
        if (n - off < 9) return SWC_DATA_TRUNCATED;
        uint32_t id = rd32(in + off); off += 4;
        if (has_ext_id && ext_id != id) return SWC_DATA_CORRUPTED;
    }
    uint32_t hc = swco_xxh32(in, off);                                            /* :272-275 */
    if ((uint8_t)((hc >> 8) & 0xFF) != in[off++]) return SWC_DATA_CORRUPTED;

    const size_t fstart = out->len;
    for (;;) {                                                                    /* :278-318 */
        if (n - off < 4) return SWC_DATA_TRUNCATED;
        uint32_t mark = rd32(in + off); off += 4;
        if (mark == 0) break;
        int compressed = (mark & 0x80000000u) == 0;
        size_t bs = mark & 0x7FFFFFFFu;
        if (bs > max_block) return SWC_DATA_CORRUPTED;
        if (n - off < bs + (block_ck ? 4 : 0) + 4) return SWC_DATA_TRUNCATED;
        const uint8_t *blk = in + off;
        off += bs;
        if (block_ck) { if (swco_xxh32(blk, bs) != rd32(in + off)) return SWC_DATA_CORRUPTED; off += 4; }
        if (compressed) {
            int st;
            if (independent) {
                st = swco_lz4_block(blk, bs, dict, have_dict ? dict_len : 0, out);               /* :305 */
            } else if (out->len == fstart && have_dict) {                                         /* :307-310 */
                size_t dl = dict_len > 65536 ? 65536 : dict_len;
                st = swco_lz4_block(blk, bs, dict + (dict_len - dl), dl, out);
            } else {                                                                              /* :311-313 */
                size_t have = out->len - fstart, dl = have > 65536 ? 65536 : have;
                /* the window aliases `out`, which may be reallocated while appending: copy it first */
                uint8_t *win = (uint8_t *)malloc(dl ? dl : 1);
                if (!win) return SWC_ERR_OUTPUT_OVERFLOW;
                memcpy(win, out->data + out->len - dl, dl);
                st = swco_lz4_block(blk, bs, win, dl, out);
                free(win);
            }
            if (st) return st;
        } else {
            if (swco_buf_append(out, blk, bs)) return SWC_ERR_OUTPUT_OVERFLOW;
        }
    }
    if (csize_p && (uint64_t)(out->len - fstart) != content_size) return SWC_DATA_CORRUPTED;
    if (cck) {                                                                    /* :323-328 */
        if (n - off < 4) return SWC_DATA_TRUNCATED;
        uint32_t stored = rd32(in + off); off += 4;
        if (swco_xxh32(out->data + fstart, out->len - fstart) != stored) { *used = off; return SWC_DATA_CHECKSUM_MISMATCH; }
    }
    *used = off;
    return SWC_OK;
}

/* decompress(data:dictionary:dictionaryID:) :73-91. dict == NULL means "no dictionary" (nil). */
int swco_lz4_decompress(const uint8_t *in, size_t n, const uint8_t *dict, size_t dict_len,
                        int has_dict_id, uint32_t dict_id, swco_buf *out, size_t *consumed) {
    size_t base = 0;
    int have_dict = dict != NULL;
    for (;;) {
        if (n - base < 4) return SWC_DATA_TRUNCATED;
        uint32_t magic = rd32(in + base);
        size_t used = 0;
        int st;
        if (magic == 0x184D2204u) {
            st = frame(in + base + 4, n - base - 4, dict, dict_len, have_dict, has_dict_id, dict_id, out, &used);
        } else if (magic >= 0x184D2A50u && magic <= 0x184D2A5Fu) {                /* :148-155, then recurse w/o dict (:85) */
            if (n - base - 4 < 4) return SWC_DATA_TRUNCATED;
            size_t size = rd32(in + base + 4);
            if (n - base - 4 < size + 4) return SWC_DATA_TRUNCATED;
            base += 4 + size + 4;
            have_dict = 0; has_dict_id = 0;
            continue;
        } else if (magic == 0x184C2102u) {
            st = legacy_frame(in + base + 4, n - base - 4, out, &used);
        } else {
            return SWC_DATA_CORRUPTED;
        }
        if (consumed) *consumed = base + 4 + used;
        return st;
    }
}

/* multiDecompress :116-146 */
int swco_lz4_multi_decompress(const uint8_t *in, size_t n, const uint8_t *dict, size_t dict_len,
                              int has_dict_id, uint32_t dict_id, swco_buf *out, size_t *ends, size_t max_frames, size_t *n_frames) {
    size_t next = 0, cnt = 0;
    do {
        if (next + 4 > n) return SWC_DATA_TRUNCATED;
        uint32_t magic = rd32(in + next); next += 4;
        size_t used = 0;
        int st, produced = 0;
        if (magic == 0x184D2204u) {
            st = frame(in + next, n - next, dict, dict_len, dict != NULL, has_dict_id, dict_id, out, &used); produced = 1;
        } else if (magic >= 0x184D2A50u && magic <= 0x184D2A5Fu) {
            if (n - next < 4) return SWC_DATA_TRUNCATED;
            size_t size = rd32(in + next);
            if (n - next < size + 4) return SWC_DATA_TRUNCATED;
            used = size + 4; st = SWC_OK;
        } else if (magic == 0x184C2102u) {
            st = legacy_frame(in + next, n - next, out, &used); produced = 1;
        } else {
            return SWC_DATA_CORRUPTED;
        }
        if (produced && (st == SWC_OK || st == SWC_DATA_CHECKSUM_MISMATCH)) {
            if (cnt < max_frames) ends[cnt] = out->len;
            cnt++;
        }
        if (n_frames) *n_frames = cnt;
        if (st) return st;
        next += used;
    } while (next < n);
    return SWC_OK;
}
