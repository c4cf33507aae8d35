/* huffman.c — ORACLE (test infrastructure): canonical code assignment + implicit-heap decoding tree.
 * Follows Sources/Common/CodingTree/Code.swift:15-39, CodeLength.swift:15-21, DecodingTree.swift:15-50 and
 * Sources/Common/Extensions.swift:43-55 (Int.reversed(bits:)).  No validation of the Kraft sum, exactly like the
 * reference: over-subscribed sets overwrite heap slots, shorter prefixes win while decoding. */
#include "swco.h"

/* Int.reversed(bits:) — Extensions.swift:43-55: reverses the low `count` bits, higher bits are dropped. */
static int64_t reversed_bits(int64_t v, int count) {
    int64_t z = 0;
    for (int i = 0; i < count; i++)
        if ((v >> i) & 1) z |= (int64_t)1 << (count - 1 - i);
    return z;
}

static int tree_alloc(swco_tree *t, int max_bits) {
    t->leaf_count = ((int64_t)1 << (max_bits + 1)) - 1;
    t->tree = (int32_t *)malloc(sizeof(int32_t) * (size_t)t->leaf_count);
    if (!t->tree) return -1;
    for (int64_t i = 0; i < t->leaf_count; i++) t->tree[i] = -1;
    return 0;
}

/* DecodingTree.init body, DecodingTree.swift:22-32 */
static void tree_put(swco_tree *t, int bits, int64_t code, int symbol) {
    int64_t index = 0;
    for (int i = 0; i < bits; i++) {
        index = (code & 1) == 0 ? 2 * index + 1 : 2 * index + 2;
        code >>= 1;
    }
    t->tree[index] = symbol;
}

int swco_tree_build(swco_tree *t, const int *lengths, int nsyms) {
    /* Code.huffmanCodes: sort by (codeLength, symbol); maxBits = last element's length (Code.swift:17-20).
       Code lengths are small (<= 20 for bzip2, <= 15 for deflate), so a counting order is the same as sorted(). */
    int max_bits = 0;
    for (int i = 0; i < nsyms; i++) if (lengths[i] > max_bits) max_bits = lengths[i];
    if (tree_alloc(t, max_bits)) return -1;
    int loop_bits = -1;
    int64_t symbol = -1;
    for (int len = 1; len <= max_bits; len++) {
        for (int s = 0; s < nsyms; s++) {
            if (lengths[s] != len) continue;
            symbol += 1;                                   /* Code.swift:27 */
            if (len != loop_bits) {                        /* Code.swift:30-33 */
                symbol <<= (len - loop_bits);
                loop_bits = len;
            }
            tree_put(t, len, reversed_bits(symbol, loop_bits), s);   /* Code.swift:35-36 + DecodingTree.swift:22-32 */
        }
    }
    return 0;
}

int swco_tree_build_codes(swco_tree *t, const int *bits, const int *codes, const int *symbols, int n, int max_bits) {
    if (tree_alloc(t, max_bits)) return -1;
    for (int i = 0; i < n; i++) tree_put(t, bits[i], codes[i], symbols[i]);
    return 0;
}

void swco_tree_free(swco_tree *t) { free(t->tree); t->tree = NULL; t->leaf_count = 0; }

/* DecodingTree.findNextSymbol, DecodingTree.swift:36-50 */
int swco_tree_next(const swco_tree *t, swco_bits *r) {
    uint64_t bits_left = swco_bits_left(r);
    int64_t index = 0;
    while (bits_left > 0) {
        unsigned bit = swco_bit(r);
        index = bit == 0 ? 2 * index + 1 : 2 * index + 2;
        bits_left -= 1;
        if (index >= t->leaf_count) return -1;
        if (t->tree[index] > -1) return t->tree[index];
    }
    return -1;
}
