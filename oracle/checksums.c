/* checksums.c — ORACLE (test infrastructure): Sources/Common/CheckSums.swift:12-57, Sources/LZ4/XxHash32.swift:24-83,
 * Sources/XZ/Sha256.swift.  The reference stores precomputed tables (CheckSums.swift:61-183); they are the standard
 * tables of the polynomials below, generated here at first use. */
#include "swco.h"

static uint32_t crc32_tab[256], bz_tab[256];
static uint64_t crc64_tab[256];
static int tabs_ready = 0;

static void make_tabs(void) {
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;      /* reflected 0x04C11DB7 */
        crc32_tab[i] = c;
        uint32_t b = i << 24;
        for (int k = 0; k < 8; k++) b = (b & 0x80000000u) ? (b << 1) ^ 0x04C11DB7u : b << 1;  /* MSB-first */
        bz_tab[i] = b;
        uint64_t d = i;
        for (int k = 0; k < 8; k++) d = (d & 1) ? 0xC96C5795D7870F42ull ^ (d >> 1) : d >> 1;  /* CRC-64/XZ */
        crc64_tab[i] = d;
    }
    tabs_ready = 1;
}

uint32_t swco_crc32(const uint8_t *p, size_t n, uint32_t prev) {          /* CheckSums.swift:12-28 */
    if (!tabs_ready) make_tabs();
    uint32_t crc = ~prev;
    for (size_t i = 0; i < n; i++) crc = crc32_tab[(crc & 0xFF) ^ p[i]] ^ (crc >> 8);
    return ~crc;
}

uint32_t swco_bzip2_crc32(const uint8_t *p, size_t n) {                   /* CheckSums.swift:30-37 */
    if (!tabs_ready) make_tabs();
    uint32_t crc = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; i++) crc = (crc << 8) ^ bz_tab[((crc >> 24) ^ p[i]) & 0xFF];
    return ~crc;
}

uint64_t swco_crc64(const uint8_t *p, size_t n) {                         /* CheckSums.swift:39-46 */
    if (!tabs_ready) make_tabs();
    uint64_t crc = ~(uint64_t)0;
    for (size_t i = 0; i < n; i++) crc = crc64_tab[(crc & 0xFF) ^ p[i]] ^ (crc >> 8);
    return ~crc;
}

uint32_t swco_adler32(const uint8_t *p, size_t n) {                       /* CheckSums.swift:48-57 */
    uint32_t s1 = 1, s2 = 0;
    for (size_t i = 0; i < n; i++) { s1 = (s1 + p[i]) % 65521u; s2 = (s2 + s1) % 65521u; }
    return (s2 << 16) + s1;
}

/* ---- xxHash32, seed 0 (XxHash32.swift:24-83) ---- */
#define P1 0x9E3779B1u
#define P2 0x85EBCA77u
#define P3 0xC2B2AE3Du
#define P4 0x27D4EB2Fu
#define P5 0x165667B1u
static inline uint32_t rotl(uint32_t v, int s) { return (v << s) | (v >> (32 - s)); }
static inline uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }

uint32_t swco_xxh32(const uint8_t *p, size_t n) {
    size_t i = 0;
    uint32_t acc;
    if (n < 16) {
        acc = P5;                                                          /* hashSmall */
    } else {
        uint32_t a[4] = {P1 + P2, P2, 0, 0u - P1};                         /* hashBig */
        while (n - i >= 16) {
            for (int j = 0; j < 4; j++) { a[j] += rd32(p + i + 4 * j) * P2; a[j] = rotl(a[j], 13); a[j] *= P1; }
            i += 16;
        }
        acc = rotl(a[0], 1) + rotl(a[1], 7) + rotl(a[2], 12) + rotl(a[3], 18);
    }
    acc += (uint32_t)n;                                                    /* finalize */
    while (n - i >= 4) { acc += rd32(p + i) * P3; acc = rotl(acc, 17) * P4; i += 4; }
    while (n - i >= 1) { acc += (uint32_t)p[i] * P5; acc = rotl(acc, 11) * P1; i += 1; }
    acc ^= acc >> 15; acc *= P2; acc ^= acc >> 13; acc *= P3; acc ^= acc >> 16;
    return acc;
}

/* ---- SHA-256 (FIPS 180-4), used by XZ check type 0x0A (Sources/XZ/Sha256.swift) ---- */
static const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
static inline uint32_t rotr(uint32_t v, int s) { return (v >> s) | (v << (32 - s)); }

static void sha256_block(uint32_t h[8], const uint8_t *b) {
    uint32_t w[64];
    for (int i = 0; i < 16; i++) w[i] = (uint32_t)b[4 * i] << 24 | (uint32_t)b[4 * i + 1] << 16 | (uint32_t)b[4 * i + 2] << 8 | b[4 * i + 3];
    for (int i = 16; i < 64; i++) {
        uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
        uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], bb = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; i++) {
        uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25), ch = (e & f) ^ (~e & g);
        uint32_t t1 = hh + S1 + ch + K256[i] + w[i];
        uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22), mj = (a & bb) ^ (a & c) ^ (bb & c);
        uint32_t t2 = S0 + mj;
        hh = g; g = f; f = e; e = d + t1; d = c; c = bb; bb = a; a = t1 + t2;
    }
    h[0] += a; h[1] += bb; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

void swco_sha256(const uint8_t *p, size_t n, uint8_t digest[32]) {
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    size_t i = 0;
    for (; i + 64 <= n; i += 64) sha256_block(h, p + i);
    uint8_t tail[128] = {0};
    size_t rem = n - i;
    memcpy(tail, p + i, rem);
    tail[rem] = 0x80;
    size_t tl = rem + 1 + 8 <= 64 ? 64 : 128;
    uint64_t bits = (uint64_t)n * 8;
    for (int k = 0; k < 8; k++) tail[tl - 1 - k] = (uint8_t)(bits >> (8 * k));
    sha256_block(h, tail);
    if (tl == 128) sha256_block(h, tail + 64);
    for (int k = 0; k < 8; k++) { digest[4 * k] = (uint8_t)(h[k] >> 24); digest[4 * k + 1] = (uint8_t)(h[k] >> 16); digest[4 * k + 2] = (uint8_t)(h[k] >> 8); digest[4 * k + 3] = (uint8_t)h[k]; }
}
