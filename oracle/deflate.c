/* deflate.c — ORACLE (test infrastructure): restatement of Sources/Deflate/Deflate.swift:30-249.
 * Line references are to that file unless another file is named. */
#include "swco.h"

/* Deflate+Constants.swift:175-186 */
static const int code_length_orders[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
static const int length_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35,
                                    43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const int distance_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193,
                                      257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145,
                                      8193, 12289, 16385, 24577};

/* Deflate+Constants.swift:11-173: the precomputed static tables are the canonical codes of the RFC 1951 fixed
 * lengths, *including* lit/len symbols 286/287 and distance symbols 30/31 (decoding them yields wrongSymbol). */
static int build_static(swco_tree *lit, swco_tree *dist) {
    int l[288], d[32];
    for (int i = 0; i < 144; i++) l[i] = 8;
    for (int i = 144; i < 256; i++) l[i] = 9;
    for (int i = 256; i < 280; i++) l[i] = 7;
    for (int i = 280; i < 288; i++) l[i] = 8;
    for (int i = 0; i < 32; i++) d[i] = 5;
    if (swco_tree_build(lit, l, 288)) return -1;
    if (swco_tree_build(dist, d, 32)) { swco_tree_free(lit); return -1; }
    return 0;
}

int swco_deflate_decompress(const uint8_t *in, size_t in_len, uint64_t start_bit, swco_buf *out, uint64_t *consumed_bits) {
    swco_bits r;
    swco_bits_init(&r, in, in_len, start_bit, 1);
    const size_t out_start = out->len;    /* `out` of this call begins here (wrappers append members) */
    int status = SWC_OK;
    swco_tree lit = {0, 0}, dist = {0, 0}, cl = {0, 0};
#define FAIL(code) do { status = (code); goto done; } while (0)

    if (swco_bits_left(&r) < 10) FAIL(SWC_DEFLATE_WRONG_BLOCK_TYPE);                      /* :36-37 */

    for (;;) {
        /* :41-43 read BFINAL/BTYPE with no bitsLeft guard: BitByteData traps when < 3 bits remain */
        if (swco_bits_left(&r) < 3) FAIL(SWC_ERR_REFERENCE_TRAP);
        unsigned is_last = swco_bit(&r);
        unsigned btype = (unsigned)swco_bits_int(&r, 2);

        if (btype == 0) {                                                                 /* :45-65 */
            swco_bits_align(&r);
            if (swco_bits_bytes_left(&r) < 4) FAIL(SWC_DEFLATE_WRONG_UNCOMPRESSED_BLOCK_LENGTHS);
            unsigned length = swco_bits_byte(&r); length |= (unsigned)swco_bits_byte(&r) << 8;
            unsigned nlength = swco_bits_byte(&r); nlength |= (unsigned)swco_bits_byte(&r) << 8;
            if ((length & nlength) != 0) FAIL(SWC_DEFLATE_WRONG_UNCOMPRESSED_BLOCK_LENGTHS); /* :56 AND, not ~ */
            if (swco_bits_bytes_left(&r) < length) FAIL(SWC_DEFLATE_WRONG_UNCOMPRESSED_BLOCK_LENGTHS);
            if (swco_buf_append(out, in + swco_bits_byte_offset(&r), length)) FAIL(SWC_ERR_OUTPUT_OVERFLOW);
            r.pos += (uint64_t)length * 8;
        } else if (btype == 1 || btype == 2) {
            if (btype == 1) {                                                             /* :77-81 */
                if (build_static(&lit, &dist)) FAIL(SWC_ERR_OUTPUT_OVERFLOW);
            } else {                                                                      /* :82-168 */
                if (swco_bits_left(&r) < 14) FAIL(SWC_DEFLATE_SYMBOL_NOT_FOUND);
                int literals = (int)swco_bits_int(&r, 5) + 257;
                if (literals > 286) FAIL(SWC_DEFLATE_WRONG_SYMBOL);                       /* :94 */
                int distances = (int)swco_bits_int(&r, 5) + 1;                            /* no upper check (:97) */
                int cl_count = (int)swco_bits_int(&r, 4) + 4;
                if (swco_bits_left(&r) < (uint64_t)(3 * cl_count)) FAIL(SWC_DEFLATE_SYMBOL_NOT_FOUND);
                int ordered[19] = {0};
                for (int i = 0; i < cl_count; i++) ordered[code_length_orders[i]] = (int)swco_bits_int(&r, 3);
                if (swco_tree_build(&cl, ordered, 19)) FAIL(SWC_ERR_OUTPUT_OVERFLOW);

                int lens[286 + 32] = {0};
                int count = literals + distances;
                int n = 0;
                while (n < count) {                                                       /* :119-158 */
                    int symbol = swco_tree_next(&cl, &r);
                    if (symbol == -1) FAIL(SWC_DEFLATE_SYMBOL_NOT_FOUND);
                    if (symbol >= 0 && symbol <= 15) {
                        lens[n++] = symbol;
                    } else if (symbol == 16 && n > 0) {
                        if (swco_bits_left(&r) < 2) FAIL(SWC_DEFLATE_SYMBOL_NOT_FOUND);
                        int copy = (int)swco_bits_int(&r, 2) + 3;
                        if (n + copy > count) FAIL(SWC_DEFLATE_WRONG_SYMBOL);
                        for (int i = 0; i < copy; i++) lens[n + i] = lens[n - 1];
                        n += copy;
                    } else if (symbol == 17) {
                        if (swco_bits_left(&r) < 3) FAIL(SWC_DEFLATE_SYMBOL_NOT_FOUND);
                        n += (int)swco_bits_int(&r, 3) + 3;
                    } else if (symbol == 18) {
                        if (swco_bits_left(&r) < 7) FAIL(SWC_DEFLATE_SYMBOL_NOT_FOUND);
                        n += (int)swco_bits_int(&r, 7) + 11;
                    } else {
                        FAIL(SWC_DEFLATE_WRONG_SYMBOL);                                   /* 16 first, or > 18 */
                    }
                }
                if (n != count) FAIL(SWC_DEFLATE_WRONG_SYMBOL);                           /* :161 */
                swco_tree_free(&cl);
                if (swco_tree_build(&lit, lens, literals)) FAIL(SWC_ERR_OUTPUT_OVERFLOW);
                if (swco_tree_build(&dist, lens + literals, distances)) FAIL(SWC_ERR_OUTPUT_OVERFLOW);
            }

            for (;;) {                                                                    /* :171-236 */
                int sym = swco_tree_next(&lit, &r);
                if (sym == -1) FAIL(SWC_DEFLATE_SYMBOL_NOT_FOUND);
                if (sym <= 255) {
                    if (swco_buf_push(out, (uint8_t)sym)) FAIL(SWC_ERR_OUTPUT_OVERFLOW);
                } else if (sym == 256) {
                    break;
                } else if (sym <= 285) {
                    int extra_len = (sym <= 260 || sym == 285) ? 0 : (((sym - 257) >> 2) - 1);
                    if (swco_bits_left(&r) < (uint64_t)extra_len) FAIL(SWC_DEFLATE_SYMBOL_NOT_FOUND);
                    int length = length_base[sym - 257] + (int)swco_bits_int(&r, extra_len);
                    int dcode = swco_tree_next(&dist, &r);
                    if (dcode == -1) FAIL(SWC_DEFLATE_SYMBOL_NOT_FOUND);
                    if (dcode > 29) FAIL(SWC_DEFLATE_WRONG_SYMBOL);
                    int extra_dist = (dcode == 0 || dcode == 1) ? 0 : ((dcode >> 1) - 1);
                    if (swco_bits_left(&r) < (uint64_t)extra_dist) FAIL(SWC_DEFLATE_SYMBOL_NOT_FOUND);
                    size_t distance = (size_t)distance_base[dcode] + (size_t)swco_bits_int(&r, extra_dist);
                    /* :214-232 — repeat the last `distance` bytes; `out` is the array of THIS call only, so a
                       distance reaching before its start is a negative array index = trap in the reference. */
                    if (distance > out->len - out_start) FAIL(SWC_ERR_REFERENCE_TRAP);
                    if (swco_buf_reserve(out, (size_t)length)) FAIL(SWC_ERR_OUTPUT_OVERFLOW);
                    size_t src = out->len - distance;
                    for (int i = 0; i < length; i++) out->data[out->len + i] = out->data[src + i];
                    out->len += (size_t)length;
                } else {
                    FAIL(SWC_DEFLATE_WRONG_SYMBOL);                                       /* 286/287 */
                }
            }
            swco_tree_free(&lit);
            swco_tree_free(&dist);
        } else {
            FAIL(SWC_DEFLATE_WRONG_BLOCK_TYPE);                                           /* :239 */
        }
        if (is_last == 1) break;
    }
done:
    swco_tree_free(&lit); swco_tree_free(&dist); swco_tree_free(&cl);
    if (consumed_bits) *consumed_bits = r.pos - start_bit;
    return status;
#undef FAIL
}
