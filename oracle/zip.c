/* zip.c — ORACLE (test infrastructure): restatement of the ZIP container walk of the reference.
 *   ZipContainer.open / info / infoWithHelper / getEntryData   Sources/ZIP/ZipContainer.swift:43-180
 *   ZipEndOfCentralDirectory.init                              Sources/ZIP/ZipEndOfCentralDirectory.swift:22-110
 *   ZipCentralDirectoryEntry.init                              Sources/ZIP/ZipCentralDirectoryEntry.swift:45-141
 *   ZipLocalHeader.init / validate                             Sources/ZIP/ZipLocalHeader.swift:39-139
 *   ZipEntryInfoHelper.init, ZipEntryInfo.init (size/type/crc) Sources/ZIP/ZipEntryInfoHelper.swift:22-44, ZipEntryInfo.swift:106-131
 *   built-in extra fields (their effect on the read offset)    Sources/ZIP/BuiltinExtraFields.swift:19-127
 *   zipString / needsUtf8                                      Sources/ZIP/LittleEndianByteReader+Zip.swift:11-100
 * Every read of the reference is an unguarded BitByteData read: out of bounds = precondition failure = SWC_ERR_REFERENCE_TRAP.
 * File names / comments are returned as byte ranges; wrongTextField is raised where String(data:encoding:.utf8) would fail
 * (CP437 decoding, taken when the UTF-8 flag is clear and the bytes do not "need" UTF-8, cannot fail). */
#include "swco.h"

typedef struct {
    const uint8_t *p; size_t n; int64_t off; int trap;
} zr;

static uint64_t rd(zr *r, int nbytes) {              /* LittleEndianByteReader.uintN / int(fromBytes:) */
    if (r->off < 0 || (uint64_t)r->off + (uint64_t)nbytes > r->n) { r->trap = 1; r->off += nbytes; return 0; }
    uint64_t v = 0;
    for (int i = 0; i < nbytes; i++) v |= (uint64_t)r->p[r->off + i] << (8 * i);
    r->off += nbytes;
    return v;
}

/* needsUtf8(), LittleEndianByteReader+Zip.swift:44-98 */
static int needs_utf8(const uint8_t *s, size_t n) {
    if (n >= 3 && s[0] == 0xEF && s[1] == 0xBB && s[2] == 0xBF) return 1;
    size_t i = 0;
    while (i < n) {
        const uint8_t b = s[i];
        if (b <= 0x7F) { i++; continue; }
        int len;
        if (b >= 0xC2 && b <= 0xDF) len = 2;
        else if (b >= 0xE0 && b <= 0xEF) len = 3;
        else if (b >= 0xF0 && b <= 0xF4) len = 4;
        else return 0;
        if (i + (size_t)len - 1 >= n) return 0;
        for (int k = 1; k < len; k++) if ((s[i + k] & 0xC0) != 0x80) return 0;
        if (len == 3) {
            const uint32_t ch = ((uint32_t)(s[i] & 0x0F) << 12) + ((uint32_t)(s[i + 1] & 0x3F) << 6) + (uint32_t)(s[i + 2] & 0x3F);
            if (ch < 0x0800 || (ch >> 11) == 0x1B) return 0;
        } else if (len == 4) {
            const uint32_t ch = ((uint32_t)(s[i] & 0x07) << 18) + ((uint32_t)(s[i + 1] & 0x3F) << 12) + ((uint32_t)(s[i + 2] & 0x3F) << 6) + (uint32_t)(s[i + 3] & 0x3F);
            if (ch < 0x10000 || ch > 0x10FFFF) return 0;
        }
        return 1;                                     /* the first multi-byte sequence decides (:96) */
    }
    return 0;
}

static int valid_utf8(const uint8_t *s, size_t n) {   /* String(data:encoding:.utf8) != nil */
    size_t i = 0;
    while (i < n) {
        const uint8_t b = s[i];
        if (b <= 0x7F) { i++; continue; }
        int len; uint32_t cp;
        if (b >= 0xC2 && b <= 0xDF) { len = 2; cp = b & 0x1F; }
        else if (b >= 0xE0 && b <= 0xEF) { len = 3; cp = b & 0x0F; }
        else if (b >= 0xF0 && b <= 0xF4) { len = 4; cp = b & 0x07; }
        else return 0;
        if (i + (size_t)len > n) return 0;
        for (int k = 1; k < len; k++) { if ((s[i + k] & 0xC0) != 0x80) return 0; cp = (cp << 6) | (s[i + k] & 0x3F); }
        if (len == 3 && (cp < 0x800 || (cp >= 0xD800 && cp <= 0xDFFF))) return 0;
        if (len == 4 && (cp < 0x10000 || cp > 0x10FFFF)) return 0;
        i += (size_t)len;
    }
    return 1;
}

/* zipString, LittleEndianByteReader+Zip.swift:11-24: returns 0 ok, 1 = nil (-> wrongTextField) */
static int zip_string(zr *r, int64_t length, int use_utf8, uint64_t *off, uint64_t *len) {
    *off = 0; *len = 0;
    if (length <= 0) return 0;
    if (r->off < 0 || (uint64_t)r->off + (uint64_t)length > r->n) { r->trap = 1; r->off += length; return 0; }
    const uint8_t *s = r->p + r->off;
    *off = (uint64_t)r->off; *len = (uint64_t)length;
    r->off += length;
    if (use_utf8) return valid_utf8(s, (size_t)length) ? 0 : 1;
    if (!needs_utf8(s, (size_t)length)) return 0;     /* CP437: every byte string decodes */
    return valid_utf8(s, (size_t)length) ? 0 : 1;
}

typedef struct {
    uint16_t version_made_by, version_needed, flags, method, time, date, internal_attrs;
    uint32_t crc, external_attrs, disk_start;
    uint64_t comp, uncomp, local_off;
    uint64_t name_off, name_len, comment_off, comment_len;
    int utf8;
    int64_t next_off;
} cd_entry;

typedef struct {
    uint16_t version_needed, flags, method, time, date;
    uint32_t crc;
    uint64_t comp, uncomp;
    int zip64;
    int64_t data_off;
} local_hdr;

/* the effect of one extra field on the reader (and, for Zip64, on the sizes) */
static void extra_field(zr *r, uint16_t id, int64_t size, int central, cd_entry *cd, local_hdr *lh) {
    switch (id) {
    case 0x0001:
        if (central) {                                                     /* ZipCentralDirectoryEntry.swift:96-107 */
            if (cd->uncomp == 0xFFFFFFFFull) cd->uncomp = rd(r, 8);
            if (cd->comp == 0xFFFFFFFFull) cd->comp = rd(r, 8);
            if (cd->local_off == 0xFFFFFFFFull) cd->local_off = rd(r, 8);
            if (cd->disk_start == 0xFFFF) cd->disk_start = (uint32_t)rd(r, 4);
        } else {                                                           /* ZipLocalHeader.swift:79-84 */
            lh->uncomp = rd(r, 8); lh->comp = rd(r, 8); lh->zip64 = 1;
        }
        break;
    case 0x5455: {                                                         /* BuiltinExtraFields.swift:19-45 */
        const int64_t end = r->off + size;
        const uint8_t flags = (uint8_t)rd(r, 1);
        if (flags & 1) rd(r, 4);
        if (!central) { if (flags & 2) rd(r, 4); if (flags & 4) rd(r, 4); }
        r->off = end;
        break;
    }
    case 0x000a: {                                                         /* :59-70 (size is ignored) */
        r->off += 4;
        const uint16_t tag = (uint16_t)rd(r, 2);
        r->off += 2;
        if (tag == 0x0001) { rd(r, 8); rd(r, 8); rd(r, 8); }
        break;
    }
    case 0x7855:
        if (central) r->off += size;                                        /* ZipCentralDirectoryEntry.swift:112-115 */
        else { rd(r, 2); rd(r, 2); }                                        /* BuiltinExtraFields.swift:84-94 */
        break;
    case 0x7875: {                                                         /* :108-127 */
        if (rd(r, 1) != 1) break;
        const int64_t us = (int64_t)rd(r, 1);
        if (us > 8) r->off += us; else rd(r, (int)us);
        const int64_t gs = (int64_t)rd(r, 1);
        if (gs > 8) r->off += gs; else rd(r, (int)gs);
        break;
    }
    default:
        r->off += size;                                                     /* no custom extra fields registered */
    }
}

static int read_cd_entry(zr *r, cd_entry *e) {                             /* ZipCentralDirectoryEntry.swift:45-141 */
    if (rd(r, 4) != 0x02014b50u) return r->trap ? SWC_ERR_REFERENCE_TRAP : SWC_ZIP_WRONG_SIGNATURE;
    e->version_made_by = (uint16_t)rd(r, 2); e->version_needed = (uint16_t)rd(r, 2);
    e->flags = (uint16_t)rd(r, 2);
    e->utf8 = (e->flags & 0x800) != 0;
    e->method = (uint16_t)rd(r, 2);
    e->time = (uint16_t)rd(r, 2); e->date = (uint16_t)rd(r, 2);
    e->crc = (uint32_t)rd(r, 4);
    e->comp = rd(r, 4); e->uncomp = rd(r, 4);
    const int64_t name_len = (int64_t)rd(r, 2), extra_len = (int64_t)rd(r, 2), comment_len = (int64_t)rd(r, 2);
    e->disk_start = (uint32_t)rd(r, 2);
    e->internal_attrs = (uint16_t)rd(r, 2); e->external_attrs = (uint32_t)rd(r, 4);
    e->local_off = rd(r, 4);
    if (r->trap) return SWC_ERR_REFERENCE_TRAP;
    if (zip_string(r, name_len, e->utf8, &e->name_off, &e->name_len)) return SWC_ZIP_WRONG_TEXT_FIELD;
    if (r->trap) return SWC_ERR_REFERENCE_TRAP;
    const int64_t start = r->off;
    while (r->off - start < extra_len) {
        const uint16_t id = (uint16_t)rd(r, 2);
        const int64_t size = (int64_t)rd(r, 2);
        if (r->trap) return SWC_ERR_REFERENCE_TRAP;
        extra_field(r, id, size, 1, e, NULL);
        if (r->trap) return SWC_ERR_REFERENCE_TRAP;
    }
    if (zip_string(r, comment_len, e->utf8, &e->comment_off, &e->comment_len)) return r->trap ? SWC_ERR_REFERENCE_TRAP : SWC_ZIP_WRONG_TEXT_FIELD;
    if (r->trap) return SWC_ERR_REFERENCE_TRAP;
    e->next_off = r->off;
    return SWC_OK;
}

static int read_local(zr *r, local_hdr *h) {                               /* ZipLocalHeader.swift:39-113 */
    if (rd(r, 4) != 0x04034b50u) return r->trap ? SWC_ERR_REFERENCE_TRAP : SWC_ZIP_WRONG_SIGNATURE;
    h->version_needed = (uint16_t)rd(r, 2);
    h->flags = (uint16_t)rd(r, 2);
    const int utf8 = (h->flags & 0x800) != 0;
    h->method = (uint16_t)rd(r, 2);
    h->time = (uint16_t)rd(r, 2); h->date = (uint16_t)rd(r, 2);
    h->crc = (uint32_t)rd(r, 4);
    h->comp = rd(r, 4); h->uncomp = rd(r, 4);
    h->zip64 = 0;
    const int64_t name_len = (int64_t)rd(r, 2), extra_len = (int64_t)rd(r, 2);
    if (r->trap) return SWC_ERR_REFERENCE_TRAP;
    uint64_t no, nl;
    if (zip_string(r, name_len, utf8, &no, &nl)) return SWC_ZIP_WRONG_TEXT_FIELD;
    if (r->trap) return SWC_ERR_REFERENCE_TRAP;
    const int64_t start = r->off;
    while (r->off - start < extra_len) {
        const uint16_t id = (uint16_t)rd(r, 2);
        const int64_t size = (int64_t)rd(r, 2);
        if (r->trap) return SWC_ERR_REFERENCE_TRAP;
        extra_field(r, id, size, 0, NULL, h);
        if (r->trap) return SWC_ERR_REFERENCE_TRAP;
    }
    h->data_off = r->off;
    return SWC_OK;
}

static int validate(const local_hdr *h, const cd_entry *e, uint32_t current_disk) {   /* ZipLocalHeader.swift:115-138 */
    if ((h->version_needed & 0xFF) > 63) return SWC_ZIP_WRONG_VERSION;
    if ((h->flags & 0x2000) || (h->flags & 0x40) || (h->flags & 0x01)) return SWC_ZIP_ENCRYPTION_NOT_SUPPORTED;
    if (h->flags & 0x20) return SWC_ZIP_PATCHING_NOT_SUPPORTED;
    if ((e->version_needed & 0xFF) > 63) return SWC_ZIP_WRONG_VERSION;
    if (e->disk_start != current_disk) return SWC_ZIP_MULTI_VOLUMES_NOT_SUPPORTED;
    if (h->flags != e->flags || h->method != e->method || h->time != e->time || h->date != e->date) return SWC_ZIP_WRONG_LOCAL_HEADER;
    return SWC_OK;
}

typedef struct {
    cd_entry cd; local_hdr lh;
    int has_dd;
    uint64_t comp, uncomp;
} helper;

/* ZipEndOfCentralDirectory.init, ZipEndOfCentralDirectory.swift:22-110 */
static int read_eocd(zr *r, uint32_t *current_disk, uint64_t *cd_entries, uint64_t *cd_offset) {
    uint32_t cur = (uint32_t)rd(r, 2), cd_disk = (uint32_t)rd(r, 2);
    if (r->trap) return SWC_ERR_REFERENCE_TRAP;
    if (cur != cd_disk) return SWC_ZIP_MULTI_VOLUMES_NOT_SUPPORTED;
    uint64_t on_disk = rd(r, 2), total = rd(r, 2);
    if (r->trap) return SWC_ERR_REFERENCE_TRAP;
    if (total != on_disk) return SWC_ZIP_MULTI_VOLUMES_NOT_SUPPORTED;
    uint64_t cd_size = rd(r, 4), off = rd(r, 4);
    if (r->trap) return SWC_ERR_REFERENCE_TRAP;
    if (cur == 0xFFFF || cd_disk == 0xFFFF || on_disk == 0xFFFF || total == 0xFFFF || cd_size == 0xFFFFFFFFull || off == 0xFFFFFFFFull) {
        r->off -= 20; r->off -= 20;
        if (rd(r, 4) != 0x07064b50u) return r->trap ? SWC_ERR_REFERENCE_TRAP : SWC_ZIP_WRONG_SIGNATURE;
        const uint32_t start_disk = (uint32_t)rd(r, 4);
        if (r->trap) return SWC_ERR_REFERENCE_TRAP;
        if (cur != start_disk) return SWC_ZIP_MULTI_VOLUMES_NOT_SUPPORTED;
        const int64_t end64 = (int64_t)rd(r, 8);
        const uint32_t total_disks = (uint32_t)rd(r, 4);
        if (r->trap) return SWC_ERR_REFERENCE_TRAP;
        if (total_disks != 1) return SWC_ZIP_MULTI_VOLUMES_NOT_SUPPORTED;
        r->off = end64;
        if (rd(r, 4) != 0x06064b50u) return r->trap ? SWC_ERR_REFERENCE_TRAP : SWC_ZIP_WRONG_SIGNATURE;
        rd(r, 8); rd(r, 2);
        const uint16_t needed = (uint16_t)rd(r, 2);
        if (r->trap) return SWC_ERR_REFERENCE_TRAP;
        if ((needed & 0xFF) > 63) return SWC_ZIP_WRONG_VERSION;
        cur = (uint32_t)rd(r, 4); cd_disk = (uint32_t)rd(r, 4);
        if (r->trap) return SWC_ERR_REFERENCE_TRAP;
        if (cur != cd_disk) return SWC_ZIP_MULTI_VOLUMES_NOT_SUPPORTED;
        on_disk = rd(r, 8); total = rd(r, 8);
        if (r->trap) return SWC_ERR_REFERENCE_TRAP;
        if (total != on_disk) return SWC_ZIP_MULTI_VOLUMES_NOT_SUPPORTED;
        cd_size = rd(r, 8); off = rd(r, 8);
        if (r->trap) return SWC_ERR_REFERENCE_TRAP;
    }
    (void)cd_size;
    *current_disk = cur; *cd_entries = total; *cd_offset = off;
    return SWC_OK;
}

/* ZipContainer.open(container:) (info_only = 0) / info(container:) (info_only = 1).
 * On SWC_OK or SWC_ZIP_WRONG_CRC `entries[0..*n_entries)` are valid; entry data is appended to `out`. */
int swco_zip_open(const uint8_t *in, size_t n, swco_buf *out, swco_zip_entry *entries, size_t max_entries, size_t *n_entries,
                  int info_only) {
    *n_entries = 0;
    if (n < 22) return SWC_ZIP_NOT_FOUND_CENTRAL_DIRECTORY_END;                         /* ZipContainer.swift:139-140 */
    zr r = {in, n, (int64_t)n - 22, 0};
    for (;;) {                                                                           /* :147-157 */
        if (rd(&r, 4) == 0x06054b50u) break;
        if (r.off == 4) return SWC_ZIP_NOT_FOUND_CENTRAL_DIRECTORY_END;
        r.off -= 5;
    }
    uint32_t current_disk; uint64_t cd_entries, cd_offset;
    int st = read_eocd(&r, &current_disk, &cd_entries, &cd_offset);
    if (st) return st;
    r.off = (int64_t)cd_offset;                                                          /* :165 */
    if (rd(&r, 4) == 0x08064b50u) r.off += (int64_t)rd(&r, 4); else r.off -= 4;          /* :166-170 */
    if (r.trap) return SWC_ERR_REFERENCE_TRAP;
    if (cd_entries > (uint64_t)n) return SWC_ERR_REFERENCE_TRAP;                         /* more entries than bytes: the walk must run off the end */
    helper *hs = (helper *)calloc((size_t)cd_entries + 1, sizeof(helper));
    if (!hs) return SWC_ERR_OUTPUT_OVERFLOW;
    for (uint64_t i = 0; i < cd_entries; i++) {                                          /* :172-177 + ZipEntryInfoHelper.swift:22-44 */
        helper *h = &hs[i];
        if ((st = read_cd_entry(&r, &h->cd))) { free(hs); return st; }
        r.off = (int64_t)h->cd.local_off;
        if ((st = read_local(&r, &h->lh))) { free(hs); return st; }
        if ((st = validate(&h->lh, &h->cd, current_disk))) { free(hs); return st; }
        h->has_dd = (h->lh.flags & 0x08) != 0;
        h->comp = h->has_dd ? h->cd.comp : h->lh.comp;
        h->uncomp = h->has_dd ? h->cd.uncomp : h->lh.uncomp;
        r.off = h->cd.next_off;
    }
    st = SWC_OK;
    size_t produced = 0;
    for (uint64_t i = 0; i < cd_entries && st == SWC_OK; i++) {
        const helper *h = &hs[i];
        swco_zip_entry e;
        memset(&e, 0, sizeof(e));
        e.name_off = h->cd.name_off; e.name_len = h->cd.name_len; e.comment_off = h->cd.comment_off; e.comment_len = h->cd.comment_len;
        e.utf8 = (uint8_t)h->cd.utf8;
        e.size = h->uncomp;                                                              /* ZipEntryInfo.swift:98 */
        e.crc = h->has_dd ? h->cd.crc : h->lh.crc;                                       /* :126 */
        e.method = h->lh.method;                                                         /* :122 */
        e.external_attrs = h->cd.external_attrs; e.version_made_by = h->cd.version_made_by; e.internal_attrs = h->cd.internal_attrs;
        e.dos_time = h->cd.time; e.dos_date = h->cd.date;
        /* entry type, ZipEntryInfo.swift:104-117: a Unix type in the high nibble wins; the DosAttributes option set always
         * exists, so the size/trailing-slash rule (:113) is unreachable */
        const uint32_t unix_type = (0xF0000000u & h->cd.external_attrs) >> 16;
        if (unix_type == 0x4000) e.is_directory = 1;
        else if (unix_type == 0x1000 || unix_type == 0x2000 || unix_type == 0x6000 || unix_type == 0x8000 || unix_type == 0xA000 || unix_type == 0xC000) e.is_directory = 0;
        else e.is_directory = (h->cd.external_attrs & 0x10) != 0;
        e.data_off = out->len; e.data_len = 0;
        if (!info_only && !e.is_directory) {                                             /* ZipContainer.getEntryData :62-125 */
            uint64_t uncomp = h->uncomp, comp = h->comp;
            uint32_t crc = e.crc;
            const int64_t data_off = h->lh.data_off;
            const size_t before = out->len;
            int64_t end_off = data_off;
            if (data_off < 0 || (uint64_t)data_off > n) { st = SWC_ERR_REFERENCE_TRAP; break; }
            const uint8_t *src = in + data_off;
            const size_t avail = n - (size_t)data_off;
            switch (h->lh.method) {
            case 0:
                if (uncomp > avail) { st = SWC_ERR_REFERENCE_TRAP; break; }
                if (swco_buf_append(out, src, (size_t)uncomp)) { st = SWC_ERR_OUTPUT_OVERFLOW; break; }
                end_off = data_off + (int64_t)uncomp;
                break;
            case 8: {
                uint64_t bits = 0;
                swco_buf tmp = {0, 0, 0};                                                /* Deflate's distances may not reach earlier entries */
                st = swco_deflate_decompress(src, avail, 0, &tmp, &bits);
                if (st == SWC_OK && swco_buf_append(out, tmp.data, tmp.len)) st = SWC_ERR_OUTPUT_OVERFLOW;
                free(tmp.data);
                end_off = data_off + (int64_t)((bits + 7) / 8);                          /* bitReader.align() :78-79 */
                break;
            }
            case 12: {
                uint64_t bits = 0;
                st = swco_bzip2_decompress(src, avail, 0, out, &bits);
                end_off = data_off + (int64_t)((bits + 7) / 8);
                break;
            }
            case 14: {
                if (avail < 9) { st = SWC_ERR_REFERENCE_TRAP; break; }                    /* 4 skipped + 5 property bytes, unguarded */
                const uint8_t pb = src[4];
                if (pb >= 225) { st = SWC_LZMA_WRONG_PROPERTIES; break; }                 /* LZMAProperties.swift:50 */
                const int64_t dict = (int64_t)src[5] | (int64_t)src[6] << 8 | (int64_t)src[7] << 16 | (int64_t)src[8] << 24;
                size_t used = 0;
                st = swco_lzma_decompress_raw(src + 9, avail - 9, pb % 9, (pb / 9) % 5, (pb / 9) / 5, dict, (int64_t)uncomp, out, &used);
                end_off = data_off + 9 + (int64_t)used;
                break;
            }
            default:
                st = SWC_ZIP_COMPRESSION_NOT_SUPPORTED;
            }
            if (st != SWC_OK) break;
            const uint64_t real_comp = (uint64_t)(end_off - data_off);
            if (h->has_dd) {                                                             /* :97-112 */
                zr d = {in, n, end_off, 0};
                if (rd(&d, 4) != 0x08074b50u) d.off -= 4;
                crc = (uint32_t)rd(&d, 4);
                if (h->lh.zip64) { comp = rd(&d, 8); uncomp = rd(&d, 8); }
                else { comp = rd(&d, 4); uncomp = rd(&d, 4); }
                if (d.trap) { st = SWC_ERR_REFERENCE_TRAP; break; }
            }
            const size_t got = out->len - before;
            if (!(comp == real_comp && uncomp == (uint64_t)got)) { st = SWC_ZIP_WRONG_SIZE; break; }    /* :114-115 */
            e.data_len = got;
            if (crc != swco_crc32(out->data + before, got, 0)) st = SWC_ZIP_WRONG_CRC;       /* :116; the entry is still returned */
        }
        if (*n_entries < max_entries) entries[*n_entries] = e;
        *n_entries += 1;
        produced = out->len;
    }
    (void)produced;
    free(hs);
    if (st != SWC_OK && st != SWC_ZIP_WRONG_CRC) *n_entries = 0;                          /* a thrown error returns nothing */
    return st;
}
