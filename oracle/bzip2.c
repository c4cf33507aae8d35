/* bzip2.c — ORACLE (test infrastructure): restatement of Sources/BZip2/BZip2.swift:40-270 and
 * Sources/BZip2/BurrowsWheeler.swift:29-64.  Line references are to BZip2.swift unless another file is named. */
#include "swco.h"

/* BurrowsWheeler.reverse(bytes:_:) BurrowsWheeler.swift:29-64. Returns 0, -1 OOM, -2 reference trap (pointer OOB). */
static int bwt_reverse(const uint8_t *bytes, size_t n, size_t pointer, uint8_t *res) {
    if (n == 0) return 0;
    size_t counts[256] = {0};
    for (size_t i = 0; i < n; i++) counts[bytes[i]]++;
    size_t base[256], sum = 0;
    for (int c = 0; c < 256; c++) { base[c] = sum; sum += counts[c]; }
    uint32_t *pointers = (uint32_t *)malloc(sizeof(uint32_t) * n);
    if (!pointers) return -1;
    for (size_t i = 0; i < n; i++) pointers[base[bytes[i]]++] = (uint32_t)i;
    size_t end = pointer;
    int rc = 0;
    for (size_t i = 0; i < n; i++) {
        if (end >= n) { rc = -2; break; }          /* pointers[end] out of range: array-bounds trap in Swift */
        end = pointers[end];
        res[i] = bytes[end];
    }
    free(pointers);
    return rc;
}

/* decode(_:_:) :97-270 — one block, appended to `out`. */
static int decode_block(swco_bits *r, swco_buf *out) {
    int status = SWC_OK;
    swco_tree tables[6];
    int ntab_built = 0;
    swco_buf buffer; swco_buf_init(&buffer);
    int *selectors = NULL;
    uint8_t *nt = NULL;
#define FAIL(c) do { status = (c); goto done; } while (0)
    if (swco_bits_left(r) < 41) FAIL(SWC_BZIP2_WRONG_MAGIC);                              /* :103 */
    if (swco_bit(r) != 0) FAIL(SWC_BZIP2_RANDOMIZED_BLOCK);                               /* :106-108 */
    size_t pointer = (size_t)swco_bits_int(r, 24);
    unsigned used_map = (unsigned)swco_bits_int(r, 16);
    if (swco_bits_left(r) < (uint64_t)(16 * __builtin_popcount(used_map) + 3 + 15)) FAIL(SWC_BZIP2_WRONG_MAGIC);
    uint8_t used[256]; int nused = 0;
    for (int blk = 0; blk < 16; blk++) {                                                  /* :122-137 */
        if (used_map & (0x8000u >> blk)) {
            unsigned m = (unsigned)swco_bits_int(r, 16);
            for (int s = 0; s < 16; s++) if (m & (0x8000u >> s)) used[nused++] = (uint8_t)(blk * 16 + s);
        }
    }
    int used_count = 2 + nused;
    int ntab = (int)swco_bits_int(r, 3);
    if (ntab < 2 || ntab > 6) FAIL(SWC_BZIP2_WRONG_HUFFMAN_GROUPS);                       /* :141-143 */
    int nsel = (int)swco_bits_int(r, 15);
    int mtf[6]; for (int i = 0; i < ntab; i++) mtf[i] = i;
    selectors = (int *)malloc(sizeof(int) * (size_t)(nsel ? nsel : 1));
    if (!selectors) FAIL(SWC_ERR_OUTPUT_OVERFLOW);
    uint64_t bits_left = swco_bits_left(r);
    for (int i = 0; i < nsel; i++) {                                                      /* :158-173 */
        int c = 0;
        while (bits_left > 0) { unsigned b = swco_bit(r); bits_left--; if (b == 0) break; c++; }
        if (c >= ntab) FAIL(SWC_BZIP2_WRONG_SELECTOR);
        int el = mtf[c];
        for (int k = c; k > 0; k--) mtf[k] = mtf[k - 1];
        mtf[0] = el;
        selectors[i] = el;
    }
    for (int t = 0; t < ntab; t++) {                                                      /* :177-203 */
        if (bits_left < 5) FAIL(SWC_BZIP2_WRONG_HUFFMAN_CODE_LENGTH);
        int length = (int)swco_bits_int(r, 5); bits_left -= 5;
        int lens[258];
        for (int i = 0; i < used_count; i++) {
            if (length < 0 || length > 20) FAIL(SWC_BZIP2_WRONG_HUFFMAN_CODE_LENGTH);
            while (bits_left > 0) {
                unsigned b = swco_bit(r); bits_left--;
                if (b == 0) break;
                if (bits_left == 0) FAIL(SWC_BZIP2_WRONG_HUFFMAN_CODE_LENGTH);
                length -= (int)swco_bit(r) * 2 - 1; bits_left--;
            }
            lens[i] = length;      /* NB :185 checks the length *before* the deltas, so the stored value may be -1 or 21 */
        }
        /* Code.huffmanCodes skips lengths <= 0 (`where length.codeLength > 0`); maxBits = largest length, and a
           negative length sorts first so it never becomes `last`. A length of 21 makes a 2^22-slot tree: still fine. */
        for (int i = 0; i < used_count; i++) if (lens[i] < 0) lens[i] = 0;
        /* Only the LAST symbol can end up > 20 (nothing re-checks it). The reference would then allocate a
           (1 << (len+1))-slot tree (32 MB at 21, unbounded above); the engine refuses instead. DESIGN.md §deviations */
        if (lens[used_count - 1] > 20) FAIL(SWC_ERR_REFERENCE_TRAP);
        if (swco_tree_build(&tables[t], lens, used_count)) FAIL(SWC_ERR_OUTPUT_OVERFLOW);
        ntab_built++;
    }

    if (nsel == 0) FAIL(SWC_ERR_REFERENCE_TRAP);                                          /* :206 selectors[0] */
    {
        int decoded = 0, sel_idx = 1;
        const swco_tree *table = &tables[selectors[0]];
        uint64_t run_length = 0, repeat_power = 1;
        for (;;) {                                                                        /* :212-246 */
            if (decoded >= 50) {
                if (sel_idx >= nsel) FAIL(SWC_BZIP2_WRONG_SELECTOR);
                table = &tables[selectors[sel_idx++]];
                decoded = 0;
            }
            int symbol = swco_tree_next(table, r);
            if (symbol == -1) FAIL(SWC_BZIP2_SYMBOL_NOT_FOUND);
            decoded++;
            if (symbol == 0 || symbol == 1) {                                             /* :226-230 (wrapping &+) */
                run_length += repeat_power << symbol;
                repeat_power <<= 1;
                continue;
            }
            if (run_length > 0) {                                                         /* :231-238 */
                if (nused == 0) FAIL(SWC_ERR_REFERENCE_TRAP);                             /* usedSymbols[0] on empty array */
                if (run_length > ((uint64_t)1 << 32)) FAIL(SWC_ERR_UNSUPPORTED);          /* would exhaust memory */
                if (swco_buf_reserve(&buffer, (size_t)run_length)) FAIL(SWC_ERR_OUTPUT_OVERFLOW);
                memset(buffer.data + buffer.len, used[0], (size_t)run_length);
                buffer.len += (size_t)run_length;
                run_length = 0; repeat_power = 1;
            }
            if (symbol == used_count - 1) break;                                          /* :239 */
            int idx = symbol - 1;                                                         /* :243-245 */
            if (idx >= nused) FAIL(SWC_ERR_REFERENCE_TRAP);
            uint8_t el = used[idx];
            for (int k = idx; k > 0; k--) used[k] = used[k - 1];
            used[0] = el;
            if (swco_buf_push(&buffer, el)) FAIL(SWC_ERR_OUTPUT_OVERFLOW);
        }
    }
    {
        size_t n = buffer.len;
        nt = (uint8_t *)malloc(n ? n : 1);
        if (!nt) FAIL(SWC_ERR_OUTPUT_OVERFLOW);
        int rc = bwt_reverse(buffer.data, n, pointer, nt);                                /* :248 */
        if (rc == -1) FAIL(SWC_ERR_OUTPUT_OVERFLOW);
        if (rc == -2) FAIL(SWC_ERR_REFERENCE_TRAP);
        size_t i = 0;                                                                     /* :251-267 */
        while (i < n) {
            if (n >= 4 && i < n - 4 && nt[i] == nt[i + 1] && nt[i] == nt[i + 2] && nt[i] == nt[i + 3]) {
                size_t run = (size_t)nt[i + 4] + 4;
                if (swco_buf_reserve(out, run)) FAIL(SWC_ERR_OUTPUT_OVERFLOW);
                memset(out->data + out->len, nt[i], run);
                out->len += run;
                i += 5;
            } else {
                if (swco_buf_push(out, nt[i])) FAIL(SWC_ERR_OUTPUT_OVERFLOW);
                i += 1;
            }
        }
    }
done:
    for (int t = 0; t < ntab_built; t++) swco_tree_free(&tables[t]);
    swco_buf_free(&buffer);
    free(selectors);
    free(nt);
    return status;
#undef FAIL
}

/* decompress(_: MsbBitReader) :50-95 — one stream. */
static int stream(swco_bits *r, swco_buf *out) {
    if (swco_bits_left(r) < 32) return SWC_BZIP2_WRONG_MAGIC;                             /* :53 */
    const size_t sstart = out->len;
    (void)sstart;
    unsigned b0 = swco_bits_byte(r), b1 = swco_bits_byte(r);                              /* uint16() little-endian */
    if ((b0 | (b1 << 8)) != 0x5a42) return SWC_BZIP2_WRONG_MAGIC;
    if (swco_bits_byte(r) != 104) return SWC_BZIP2_WRONG_VERSION;
    unsigned bs = swco_bits_byte(r);
    if (bs < 0x31 || bs > 0x39) return SWC_BZIP2_WRONG_BLOCK_SIZE;                        /* BZip2+BlockSize.swift:29-52 */
    uint32_t total_crc = 0;
    for (;;) {
        if (swco_bits_left(r) < 80) return SWC_BZIP2_WRONG_MAGIC;                         /* :71 */
        uint64_t block_type = swco_bits_int(r, 48);
        uint32_t block_crc = (uint32_t)swco_bits_int(r, 32);
        if (block_type == 0x314159265359ull) {
            size_t bstart = out->len;
            int st = decode_block(r, out);
            if (st) return st;
            if (swco_bzip2_crc32(out->data + bstart, out->len - bstart) != block_crc) return SWC_BZIP2_WRONG_CRC; /* :81 */
            total_crc = (total_crc << 1) | (total_crc >> 31);
            total_crc ^= block_crc;
        } else if (block_type == 0x177245385090ull) {
            if (total_crc != block_crc) return SWC_BZIP2_WRONG_CRC;                       /* :86 */
            break;
        } else {
            return SWC_BZIP2_WRONG_BLOCK_TYPE;
        }
    }
    return SWC_OK;
}

int swco_bzip2_decompress(const uint8_t *in, size_t in_len, uint64_t start_bit, swco_buf *out, uint64_t *consumed_bits) {
    swco_bits r;
    swco_bits_init(&r, in, in_len, start_bit, 0);
    int st = stream(&r, out);
    if (consumed_bits) *consumed_bits = r.pos - start_bit;
    return st;
}

/* multiDecompress :40-48 */
int swco_bzip2_multi_decompress(const uint8_t *in, size_t in_len, swco_buf *out, size_t *ends, size_t max_n, size_t *n) {
    swco_bits r;
    swco_bits_init(&r, in, in_len, 0, 0);
    size_t cnt = 0;
    if (n) *n = 0;
    while (swco_bits_byte_offset(&r) < in_len) {         /* !reader.isFinished (byte granular; reader is aligned here) */
        int st = stream(&r, out);
        if (st) return st;
        if (cnt < max_n) ends[cnt] = out->len;
        cnt++;
        if (n) *n = cnt;
        swco_bits_align(&r);
    }
    return SWC_OK;
}
