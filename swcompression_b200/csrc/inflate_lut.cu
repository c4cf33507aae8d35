// inflate_lut.cu — K1F: fused Deflate decoder for large batches (Huffman walk + LZ77 copy in ONE kernel, no records).
// Replaces Deflate.decompress(_: LsbBitReader) (reference Sources/Deflate/Deflate.swift:30-249) with
// DecodingTree.findNextSymbol (Sources/Common/CodingTree/DecodingTree.swift:36-50) for the batched hot path.
//
// ONE THREAD PER UNIT for the bitstream (32 different streams per warp, persistent lanes fed by a ticket counter) and
// THE WHOLE WARP for every byte that is written.  A warp works in rounds that start with a full-mask vote (lock step):
//   top-up : every lane keeps an 8-word ring of its compressed stream in shared memory; a lane whose ring is half empty
//            stores the 16-byte chunk it prefetched a round earlier (ld.global.nc.L1::no_allocate.v4) and issues the next
//            load — the only place global input is touched, so the load latency never sits on the decode chain.
//   fast   : up to KLIT table lookups per lane: peek (funnel shift of a 64-bit register window) -> 2^8-entry 16-bit LUT in
//            shared memory, halfword-interleaved across the warp (entry h of lane l at halfword h*32+l: at most a 2-way bank
//            conflict) -> literal: shifted into the lane's 8-byte fragment register; anything else (length, end of block,
//            code longer than 8 bits) parks the lane.  No stores, no position arithmetic on this path.
//   parked : all parked lanes together: length extra bits, distance code (2^5-entry LUT, canonical limit-compare decoder for
//            longer codes), reference checks (Deflate.swift:199-232); the match becomes the lane's PENDING match and the
//            32-byte sectors of its source are prefetched into L2.
//   write  : (a) the warp copies the matches that were pending from the PREVIOUS round — 16 at a time, four 8-lane groups x
//            four loads in flight before the first store, so one L2 round trip covers 16 matches and the DRAM latency of a far
//            source was spent during the round in between; overlapping matches replicate their period.
//            (b) the warp writes this round's literal fragments (<= 7 bytes per lane) with the same 8-lane groups.
//            Stream order per unit is  ... match(r-1) < fragment(r) < match(r) ...  and (a) runs before (b) before the next
//            round's (a), so every source byte is final when it is read.  Parameters travel through a 512-byte staging
//            area per warp (one LDS.128 per group instead of five shuffles).
//   header : block headers (Deflate.swift:41-168) are parsed by the lanes that reached one, with a plain global-memory bit
//            reader; short codes fill the LUTs, long ones go to sorted lists for the canonical decoder.
// There is no match-record stream and no second kernel: HBM sees the compressed bytes once, the output once, plus the
// 32-byte sectors of far match sources.
// Input availability (the reference's bitsLeft guards) is checked lazily against an absolute bit position: reads past the
// unit return zero bits and the first field that crosses the end reports symbolNotFound exactly as the reference does.
// Code sets with Kraft sum > 1 go to inflate_slow_kernel via SWC_INTERNAL_NEEDS_SLOW (same contract as K1).
#include "common.cuh"
#include "inflate.cuh"

namespace swc {
namespace inflate {
namespace k1f {

constexpr int LB = 8;                       // lit/len LUT index bits
constexpr int DB = 5;                       // distance LUT index bits
// ---- per-lane shared memory: halfword area (entry h of lane l at H[h*32+l]) then word area (word w at W[w*32+l]) ----
constexpr int H_LIT = 0;
constexpr int H_DST = 1 << LB;
constexpr int H_TOTAL = (1 << LB) + (1 << DB);          // 288 halfwords
constexpr int W_LIT_BO = 0;                 // [1..15] first left-justified 15-bit code | index of its first LONG symbol << 16
constexpr int W_DST_BO = 16;
constexpr int W_CL_BO = 32;                 // [1..7]
constexpr int W_CL_SYM = 40;                // 19 x u8
constexpr int W_TOTAL = 45;
constexpr int RING_BYTES = 8 * 32 * 4;      // per warp: 8 words of compressed input per lane; 1 KiB, 1 KiB-aligned (address wrap by mask)
constexpr int WARP_BYTES = H_TOTAL * 32 * 2 + W_TOTAL * 32 * 4;
constexpr int LUT_WORDS = 64;               // CTA-shared length / distance base+extra tables

#ifndef SWC_KLIT2
#define SWC_KLIT2 6
#endif
constexpr int KLIT = SWC_KLIT2;             // lookups a lane may do per round (<= 48 bits) before parked symbols are serviced

constexpr u32 E_NONLIT = 0x8000u;           // LUT entry: bit15 = not a literal; [11:8] code length (0 = long / no code)
constexpr u32 CODE_EOB = 31;                //   non-literal low byte: 0..28 length symbol 257+k, 29/30 = 286/287, 31 = end of block

// RFC 1951 3.2.5 tables as {base | extra_bits << 16}; Deflate+Constants.swift:179-186 + Deflate.swift:188-189,206
__constant__ u32 c_len_tab[32] = {
    3, 4, 5, 6, 7, 8, 9, 10, 11 | 1 << 16, 13 | 1 << 16, 15 | 1 << 16, 17 | 1 << 16, 19 | 2 << 16, 23 | 2 << 16, 27 | 2 << 16,
    31 | 2 << 16, 35 | 3 << 16, 43 | 3 << 16, 51 | 3 << 16, 59 | 3 << 16, 67 | 4 << 16, 83 | 4 << 16, 99 | 4 << 16,
    115 | 4 << 16, 131 | 5 << 16, 163 | 5 << 16, 195 | 5 << 16, 227 | 5 << 16, 258, 0, 0, 0};
__constant__ u32 c_dist_tab[32] = {
    1, 2, 3, 4, 5 | 1 << 16, 7 | 1 << 16, 9 | 2 << 16, 13 | 2 << 16, 17 | 3 << 16, 25 | 3 << 16, 33 | 4 << 16, 49 | 4 << 16,
    65 | 5 << 16, 97 | 5 << 16, 129 | 6 << 16, 193 | 6 << 16, 257 | 7 << 16, 385 | 7 << 16, 513 | 8 << 16, 769 | 8 << 16,
    1025 | 9 << 16, 1537 | 9 << 16, 2049 | 10 << 16, 3073 | 10 << 16, 4097 | 11 << 16, 6145 | 11 << 16, 8193 | 12 << 16,
    12289 | 12 << 16, 16385 | 13 << 16, 24577 | 13 << 16, 0, 0};
__constant__ u8 c_cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

struct Limits { u32 p[8]; };   // p[k] = limit[2k+1] | limit[2k+2] << 16 ; limit[L] = left-justified end of the length-L code range

__device__ __forceinline__ int code_length(u32 r15, const Limits &lim) {
    const u32 X = (r15 | (r15 << 16)) + 0x80008000u;
    u32 t[8];
#pragma unroll
    for (int k = 0; k < 8; k++) t[k] = X - lim.p[k];
    u32 a = __byte_perm(t[0], t[1], 0x7531), b = __byte_perm(t[2], t[3], 0x7531);
    u32 c = __byte_perm(t[4], t[5], 0x7531), d = __byte_perm(t[6], t[7], 0x7531);
    u32 v = (a & 0x80808080u) | ((b & 0x80808080u) >> 1) | ((c & 0x80808080u) >> 2) | ((d & 0x80808080u) >> 3);
    return 1 + __popc(v);                                    // 16 => no code matches (incomplete set)
}

// The unit as seen by both readers: bit positions count from `origin`, the 16-byte aligned address at or below the unit's
// first byte.  Bytes outside [ubeg, uend) read as zero.
struct Span {
    const u8 *origin, *ubeg, *uend;
    u32 pos0;        // bit position of the unit's first bit
    u32 end;         // pos0 + unit length in bits  (the reference's bitsLeft == end - pos)
};

__device__ __noinline__ uint4 load_edge(const u8 *c, const u8 *ubeg, const u8 *uend) {   // chunk straddling an end of the unit
    u32 w[4] = {0, 0, 0, 0};
    if (c < uend && c + 16 > ubeg) {
        for (int k = 0; k < 16; k++) {
            const u8 *a = c + k;
            if (a >= ubeg && a < uend) w[k >> 2] |= (u32)__ldg(a) << ((k & 3) * 8);
        }
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ uint4 load_chunk(const Span &sp, u32 ch) {
    const u8 *c = sp.origin + (size_t)ch * 16;
    if (c >= sp.ubeg && c + 16 <= sp.uend) {
        uint4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(c));
        return v;
    }
    return load_edge(c, sp.ubeg, sp.uend);
}

__device__ __forceinline__ u32 lds32(u32 saddr) { u32 v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr)); return v; }
__device__ __forceinline__ u32 lds16(u32 saddr) { u32 v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(saddr)); return v; }
// LUT entry address: base + index * 64 (halfword index*32 + lane), as one multiply-add
__device__ __forceinline__ u32 lut_addr(u32 index, u32 base) { u32 a; asm("mad.lo.u32 %0, %1, 64, %2;" : "=r"(a) : "r"(index), "r"(base)); return a; }

// ------------------------------------------------------------------------------------------------ symbol-phase bit reader
// Window (lo, hi) = stream words [wend/32 - 1, wend/32]; further words wait in the shared-memory ring (slot = word index & 7,
// slot k of lane l at ring + k*128 + l*4; the ring of a warp is 1 KiB-aligned so the slot address wraps with one LOP3).
struct Reader {
    u32 lo, hi;
    u32 pos;          // absolute bit position of the next unread bit;  wend - 32 <= pos < wend + 32
    u32 wend;         // bit position where `hi` starts; the next word to pop has index wend/32 + 1
    u32 rptr;         // shared-memory address of that word's slot
    u32 wr;           // next word index to push into the ring
    u32 nextc;        // chunk index after `pre`
    uint4 pre;        // chunk wr/4, already loaded

    __device__ __forceinline__ u32 peek32() const { return __funnelshift_r(lo, hi, pos); }   // requires pos < wend
    __device__ __forceinline__ void advance() {                                                // requires pos >= wend
        lo = hi;
        hi = lds32(rptr);
        const u32 t = rptr + 128;
        rptr = (t & 0x380u) | (rptr & ~0x380u);
        wend += 32;
    }
    __device__ __forceinline__ void topup(u32 *ring, const Span &sp) {
        if (wr - (wend >> 5) <= 5) {                               // <= 4 unread words in the ring: chunk wr/4 - 2 is consumed
            u32 *s = ring + (wr & 4) * 32;
            s[0] = pre.x; s[32] = pre.y; s[64] = pre.z; s[96] = pre.w;
            wr += 4;
            pre = load_chunk(sp, nextc);
            nextc++;
        }
    }
    __device__ void seek(u32 *ring, const Span &sp, u32 bit) {
        const u32 w0 = bit >> 5, ch = w0 >> 2;
        const uint4 a = load_chunk(sp, ch), b = load_chunk(sp, ch + 1);
        pre = load_chunk(sp, ch + 2);
        nextc = ch + 3;
        u32 *s = ring + (ch & 1) * 4 * 32, *t = ring + ((ch + 1) & 1) * 4 * 32;
        s[0] = a.x; s[32] = a.y; s[64] = a.z; s[96] = a.w;
        t[0] = b.x; t[32] = b.y; t[64] = b.z; t[96] = b.w;
        wr = (ch + 2) * 4;
        lo = ring[(w0 & 7) * 32];
        hi = ring[((w0 + 1) & 7) * 32];
        rptr = (u32)__cvta_generic_to_shared(ring + ((w0 + 2) & 7) * 32);
        wend = (w0 + 1) * 32;
        pos = bit;
    }
};

// ------------------------------------------------------------------------------------------------ header-phase bit reader
struct HeaderBits {
    Span sp;
    u32 pos;
    __device__ __forceinline__ u32 word(u32 i) const {
        const u8 *a = sp.origin + (size_t)i * 4;
        if (a >= sp.ubeg && a + 4 <= sp.uend) return __ldg((const u32 *)a);
        u32 v = 0;
        for (int k = 0; k < 4; k++)
            if (a + k >= sp.ubeg && a + k < sp.uend) v |= (u32)__ldg(a + k) << (8 * k);
        return v;
    }
    __device__ __forceinline__ u32 peek32() const { return __funnelshift_r(word(pos >> 5), word((pos >> 5) + 1), pos); }
    __device__ __forceinline__ u32 take(int n) { const u32 v = peek32() & ((1u << n) - 1); pos += n; return v; }
    __device__ __forceinline__ i64 avail() const { return (i64)sp.end - (i64)pos; }
};

// ------------------------------------------------------------------------------------------------ output side
struct Emit {
    u8 *out;                 // unit output base (16-byte aligned)
    u32 op;                  // bytes produced up to the start of this round's fragment
    u32 cap;
    u32 acc_lo, acc_hi;      // this round's literals: the most recent in the top byte
    u32 nf;                  // how many

    __device__ __forceinline__ void literal(u32 e) {          // low byte of e
        acc_lo = __funnelshift_r(acc_lo, acc_hi, 8);
        acc_hi = __funnelshift_r(acc_hi, e, 8);
        nf++;
    }
    __device__ __forceinline__ void stored_byte(u32 b) {      // header phase: this lane's fragment has been posted (nf == 0)
        if (op < cap) out[op] = (u8)b;
        op++;
    }
};

// ------------------------------------------------------------------------------------------------ table construction
// finalize one alphabet: bo[L] holds count[L] on entry, {first_code_lj | first_long_index << 16} on exit; `lutbits` = codes of
// at most that many bits live in the LUT and are not indexed.  Returns the Kraft sum scaled to 2^15.
__device__ __forceinline__ u32 finalize_tables(u32 *bo, Limits &lim, int maxlen, int lutbits) {
    u32 code = 0, off = 0;
    u32 l[17];
#pragma unroll
    for (int L = 1; L <= 15; L++) {
        u32 c = L <= maxlen ? bo[L * 32] : 0;
        if (L <= maxlen) bo[L * 32] = (code & 0xFFFF) | (off << 16);
        code += c << (15 - L);
        if (L > lutbits) off += c;
        l[L] = code > 0x8000u ? 0x8000u : code;
    }
    l[16] = 0x8000u;
#pragma unroll
    for (int k = 0; k < 8; k++) lim.p[k] = l[2 * k + 1] | (l[2 * k + 2] << 16);
    return code;
}
// after the assignment pass bo[L] = {end code of length L | end index}: the canonical property makes that the first
// code / index of length L+1, so shifting the array up by one slot restores the "first" values
__device__ __forceinline__ void rewind_tables(u32 *bo, int maxlen) {
    u32 prev = 0;
#pragma unroll
    for (int L = 1; L <= 15; L++) {
        if (L <= maxlen) {
            const u32 w = bo[L * 32];
            bo[L * 32] = prev;
            prev = w;
        }
    }
}

__device__ __forceinline__ int static_len(int i) {   // i < 288: lit/len, else distance (32 symbols of 5 bits)
    return i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : i < 288 ? 8 : 5;
}

struct LaneMem {
    u16 *H;        // this lane's halfword 0
    u32 *W;        // this lane's word 0
    u16 *longsym;  // lit/len symbols with codes longer than LB bits, in canonical order (local memory)
    u8 *longdst;   // distance symbols with codes longer than DB bits (local memory)
};

// code-length alphabet: canonical decode (<= 7-bit codes, 19 symbols)
__device__ __forceinline__ int decode_cl(u32 peek, const Limits &lim, const LaneMem &M, int &L) {
    const u32 r15 = __brev(peek & 0x7FFFu) >> 17;
    L = code_length(r15, lim);
    if (L > 15) return -1;
    const u32 w = M.W[(W_CL_BO + L) * 32];
    const u32 idx = (w >> 16) + ((r15 - (w & 0xFFFFu)) >> (15 - L));
    return ((const u8 *)(M.W + (W_CL_SYM + (idx >> 2)) * 32))[idx & 3];
}

// One pass over the HLIT+HDIST code lengths of a dynamic block (Deflate.swift:119-161) or the fixed lengths of a static
// block (Deflate+Constants.swift:11-173). PASS 0 counts lengths into the BO areas; PASS 1 assigns codes: short codes fill
// the LUTs, long ones are appended to the sorted symbol lists.
template <int PASS>
__device__ int run_lengths(HeaderBits &hb, const LaneMem &M, const Limits &cl_lim, bool dynamic, int hlit, int hdist) {
    const int count = hlit + hdist;
    int n = 0, prev = 0;
    while (n < count) {
        int len, reps = 1;
        if (!dynamic) {
            len = static_len(n < hlit ? n : 288 + (n - hlit));
        } else {
            int cl;
            const int sym = decode_cl(hb.peek32(), cl_lim, M, cl);
            if (sym < 0 || hb.avail() < cl) return SWC_DEFLATE_SYMBOL_NOT_FOUND;
            hb.pos += cl;
            if (sym <= 15) {
                len = sym;
            } else if (sym == 16) {
                if (n == 0) return SWC_DEFLATE_WRONG_SYMBOL;
                if (hb.avail() < 2) return SWC_DEFLATE_SYMBOL_NOT_FOUND;
                reps = (int)hb.take(2) + 3;
                if (n + reps > count) return SWC_DEFLATE_WRONG_SYMBOL;
                len = prev;
            } else if (sym == 17) {
                if (hb.avail() < 3) return SWC_DEFLATE_SYMBOL_NOT_FOUND;
                reps = (int)hb.take(3) + 3;
                len = 0;
            } else {   // 18 (the alphabet has 19 symbols)
                if (hb.avail() < 7) return SWC_DEFLATE_SYMBOL_NOT_FOUND;
                reps = (int)hb.take(7) + 11;
                len = 0;
            }
        }
        if (len == 0) {
            n += reps;      // zeros: nothing to count or place (may overshoot `count`: checked below)
        } else {
            for (int r = 0; r < reps; r++, n++) {
                const bool is_lit = n < hlit;
                u32 *bo = M.W + ((is_lit ? W_LIT_BO : W_DST_BO) + len) * 32;
                if (PASS == 0) {
                    *bo += 1;
                } else {
                    const u32 w = *bo;
                    const int lutbits = is_lit ? LB : DB;
                    const u32 sym = is_lit ? (u32)n : (u32)(n - hlit);
                    if (len <= lutbits) {
                        *bo = w + (1u << (15 - len));
                        u32 e;
                        if (!is_lit) e = ((u32)len << 8) | sym;
                        else if (sym < 256) e = ((u32)len << 8) | sym;
                        else e = E_NONLIT | ((u32)len << 8) | (sym == 256 ? CODE_EOB : sym - 257);
                        u16 *lut = M.H + (is_lit ? H_LIT : H_DST) * 32;
                        const u32 step = 1u << len, lim = 1u << lutbits;
                        for (u32 k = __brev(w & 0xFFFFu) >> 17; k < lim; k += step) lut[k * 32] = (u16)e;
                    } else {
                        *bo = w + (1u << (15 - len)) + 0x10000u;
                        const u32 pos = w >> 16;
                        if (is_lit) M.longsym[pos] = (u16)sym;
                        else M.longdst[pos] = (u8)sym;
                    }
                }
            }
        }
        prev = len;
    }
    if (n != count) return SWC_DEFLATE_WRONG_SYMBOL;          // Deflate.swift:161
    return SWC_OK;
}

enum { ST_HEADER = 0, ST_SYMBOLS = 1, ST_PARKED = 2, ST_DONE = 3 };

struct BlockCtx {
    Limits lit_lim, dst_lim;
    bool is_last;
};

// Block header (Deflate.swift:41-168): stored blocks are copied here; for Huffman blocks the per-lane tables are built.
// On entry hb.pos is the header's first bit; on SWC_OK exit it is the first bit after the header (or after the stored bytes).
__device__ __forceinline__ int begin_block(HeaderBits &hb, Emit &em, const LaneMem &M, BlockCtx &bc, int &next) {
    if (hb.avail() < 3) return SWC_ERR_REFERENCE_TRAP;                                  // :41-43 unguarded reads
    const u32 hdr = hb.take(3);
    bc.is_last = (hdr & 1) != 0;
    const u32 btype = hdr >> 1;
    if (btype == 3) return SWC_DEFLATE_WRONG_BLOCK_TYPE;                                 // :239
    if (btype == 0) {                                                                   // :45-65
        hb.pos += (u32)(hb.avail() & 7);
        if (hb.avail() < 32) return SWC_DEFLATE_WRONG_UNCOMPRESSED_BLOCK_LENGTHS;
        const u32 length = hb.take(16);
        const u32 nlength = hb.take(16);
        if ((length & nlength) != 0) return SWC_DEFLATE_WRONG_UNCOMPRESSED_BLOCK_LENGTHS;
        if ((hb.avail() >> 3) < (i64)length) return SWC_DEFLATE_WRONG_UNCOMPRESSED_BLOCK_LENGTHS;
        const u8 *src = hb.sp.origin + (hb.pos >> 3);                                   // byte aligned here
        for (u32 i = 0; i < length; i++) em.stored_byte(__ldg(src + i));
        hb.pos += length * 8;
        next = bc.is_last ? ST_DONE : ST_HEADER;
        return SWC_OK;
    }
    const bool dynamic = btype == 2;
    int hlit = 288, hdist = 32;
    Limits cl_lim;
    u32 lens_pos = hb.pos;
#pragma unroll
    for (int L = 1; L <= 15; L++) { M.W[(W_LIT_BO + L) * 32] = 0; M.W[(W_DST_BO + L) * 32] = 0; }
    if (dynamic) {
        if (hb.avail() < 14) return SWC_DEFLATE_SYMBOL_NOT_FOUND;
        hlit = (int)hb.take(5) + 257;
        if (hlit > 286) return SWC_DEFLATE_WRONG_SYMBOL;                                 // :94
        hdist = (int)hb.take(5) + 1;
        const int hclen = (int)hb.take(4) + 4;
        if (hb.avail() < 3 * hclen) return SWC_DEFLATE_SYMBOL_NOT_FOUND;
        u64 cl = 0;                                 // 19 x 3-bit code lengths, indexed by symbol
        for (int i = 0; i < hclen; i++) cl |= (u64)hb.take(3) << (3 * c_cl_order[i]);
        u64 cnt = 0;                                // 8 x 8-bit counters
        for (int s = 0; s < 19; s++) cnt += 1ull << (8 * ((cl >> (3 * s)) & 7));
#pragma unroll
        for (int L = 1; L <= 7; L++) M.W[(W_CL_BO + L) * 32] = (u32)(cnt >> (8 * L)) & 0xFF;
        const u32 kraft = finalize_tables(M.W + W_CL_BO * 32, cl_lim, 7, 0);
        if (kraft > 0x8000u) return SWC_INTERNAL_NEEDS_SLOW;
        for (int s = 0; s < 19; s++) {
            const u32 l = (u32)(cl >> (3 * s)) & 7;
            if (l) {
                const u32 w = M.W[(W_CL_BO + l) * 32];
                M.W[(W_CL_BO + l) * 32] = w + (1u << (15 - l)) + 0x10000u;
                const u32 pos = w >> 16;
                ((u8 *)(M.W + (W_CL_SYM + (pos >> 2)) * 32))[pos & 3] = (u8)s;
            }
        }
        rewind_tables(M.W + W_CL_BO * 32, 7);
        lens_pos = hb.pos;
    }
    int st = run_lengths<0>(hb, M, cl_lim, dynamic, hlit, hdist);
    if (st) return st;
    const u32 k1 = finalize_tables(M.W + W_LIT_BO * 32, bc.lit_lim, 15, LB);
    const u32 k2 = finalize_tables(M.W + W_DST_BO * 32, bc.dst_lim, 15, DB);
    if (k1 > 0x8000u || k2 > 0x8000u) return SWC_INTERNAL_NEEDS_SLOW;
    for (int k = 0; k < (1 << LB); k++) M.H[(H_LIT + k) * 32] = (u16)E_NONLIT;     // "no short code here": canonical decoder decides
    for (int k = 0; k < (1 << DB); k++) M.H[(H_DST + k) * 32] = 0;
    hb.pos = lens_pos;
    run_lengths<1>(hb, M, cl_lim, dynamic, hlit, hdist);
    rewind_tables(M.W + W_LIT_BO * 32, 15);
    rewind_tables(M.W + W_DST_BO * 32, 15);
    next = ST_SYMBOLS;
    return SWC_OK;
}

struct Match { u32 len, dist; };

// A symbol the LUT could not finish: long code, end of block, or a length + distance (Deflate.swift:171-232).
// Returns a status; on a match `mt.len` is non-zero and the emitter has NOT been advanced yet.
__device__ __forceinline__ int parked_step(Reader &br, Emit &em, const LaneMem &M, const BlockCtx &bc, const Span &sp,
                                           const u32 *lut, int &state, u32 e, Match &mt) {
    // br.pos < br.wend here: the window still holds the >= 32 bits the lookup saw
    const u32 w0 = br.peek32();                                                // lit/len code (<= 15) + extra bits (<= 5) lie in here
    u32 L = (e >> 8) & 15, code = e & 0xFF;
    if (L == 0) {                                                              // canonical decode of a long (or missing) code
        const u32 r15 = __brev(w0 & 0x7FFFu) >> 17;
        const int CL = code_length(r15, bc.lit_lim);
        if (CL > 15) return SWC_DEFLATE_SYMBOL_NOT_FOUND;
        const u32 w = M.W[(W_LIT_BO + CL) * 32];
        const u32 sym = M.longsym[(w >> 16) + ((r15 - (w & 0xFFFFu)) >> (15 - CL))];
        br.pos += CL;
        L = (u32)CL;
        if (sym < 256) {
            if (br.pos > sp.end) return SWC_DEFLATE_SYMBOL_NOT_FOUND;
            em.literal(sym);
            state = ST_SYMBOLS;
            return SWC_OK;
        }
        code = sym == 256 ? CODE_EOB : sym - 257;
    } else {
        br.pos += L;
    }
    if (br.pos > sp.end) return SWC_DEFLATE_SYMBOL_NOT_FOUND;                  // the code itself crossed the end of the input
    if (code == CODE_EOB) { state = bc.is_last ? ST_DONE : ST_HEADER; return SWC_OK; }
    if (code > 28) return SWC_DEFLATE_WRONG_SYMBOL;                            // 286 / 287
    const u32 le = lut[code];
    const u32 eb = le >> 16;
    const u32 length = (le & 0xFFFFu) + ((w0 >> L) & ((1u << eb) - 1));
    br.pos += eb;
    if (br.pos >= br.wend) br.advance();
    const u32 w32 = br.peek32();                                               // distance code (<= 15) + extra bits (<= 13)
    const u32 de = M.H[(H_DST + (w32 & ((1u << DB) - 1))) * 32];
    u32 DL = de >> 8, dsym = de & 0xFF;
    if (DL == 0) {
        const u32 r15 = __brev(w32 & 0x7FFFu) >> 17;
        DL = (u32)code_length(r15, bc.dst_lim);
        if (DL > 15) return SWC_DEFLATE_SYMBOL_NOT_FOUND;
        const u32 w = M.W[(W_DST_BO + DL) * 32];
        dsym = M.longdst[(w >> 16) + ((r15 - (w & 0xFFFFu)) >> (15 - DL))];
    }
    br.pos += DL;
    if (br.pos > sp.end) return SWC_DEFLATE_SYMBOL_NOT_FOUND;
    if (dsym > 29) return SWC_DEFLATE_WRONG_SYMBOL;
    const u32 dd = lut[32 + dsym];
    const u32 db = dd >> 16;
    const u32 dist = (dd & 0xFFFFu) + ((w32 >> DL) & ((1u << db) - 1));
    br.pos += db;
    if (br.pos > sp.end) return SWC_DEFLATE_SYMBOL_NOT_FOUND;
    if (dist > em.op + em.nf) return SWC_ERR_REFERENCE_TRAP;                             // :219 negative array index
    if ((u64)em.op + em.nf + length > 0xFFFFFFF0ull) return SWC_ERR_UNSUPPORTED;
    mt.len = length; mt.dist = dist;
    state = ST_SYMBOLS;
    return SWC_OK;
}

// ------------------------------------------------------------------------------------------------ decoder -> copier queue
// One queue per (decoder warp, copier warp) pair: QSLOTS slots of 32 match entries + 32 fragment entries (16 B each) + a
// header word, guarded by a full/empty mbarrier pair per slot (producer/consumer pipeline; phases flip every QSLOTS rounds).
//   match entry    {dst lo, dst hi, len | dist << 16, sum of the lengths of the entries before it}
//   fragment entry {dst lo, dst hi | bytes << 24, literal bytes 0-3, literal bytes 4-7}
constexpr int QSLOTS = 2;
constexpr int SLOT_BYTES = 1024 + 16;        // entries + header {n_match | n_frag << 8 | done << 16, total match bytes}
constexpr int QUEUE_BYTES = QSLOTS * SLOT_BYTES + QSLOTS * 2 * 8;   // + full[QSLOTS], empty[QSLOTS] mbarriers
constexpr u32 Q_DONE = 1u << 16;

__device__ __forceinline__ void mbar_init(u32 saddr, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(saddr), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(u32 saddr) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(saddr) : "memory");
}
__device__ __forceinline__ void mbar_wait(u32 saddr, u32 parity) {
    asm volatile(
        "{\n .reg .pred p;\n"
        "MBAR_WAIT:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra MBAR_DONE;\n bra MBAR_WAIT;\n"
        "MBAR_DONE:\n}" ::"r"(saddr), "r"(parity) : "memory");
}

__device__ __forceinline__ u8 *make_ptr(u32 lo, u32 hi) { return (u8 *)(((uintptr_t)hi << 32) | lo); }

// ---- copier warp: applies the decoder's output in stream order.  Per slot: literal fragments first (a match may read its
// own round's fragment), then ALL match bytes of the slot as one flat list: byte b of the list belongs to the entry whose
// length prefix covers it (binary search over the <= 32 prefixes in the slot), lane l takes bytes l, l+32, ... and issues all
// its loads before its first store.  The matches of a slot belong to different units, so every byte is independent: one
// memory round trip per slot whatever the match lengths.  Deflate.swift:222-229 copies byte by byte, so a match longer than
// its distance repeats its first `dist` bytes: byte i reads source byte i mod dist.
__device__ __forceinline__ void copier_loop(u8 *queue, u32 lane) {
    const u32 sub = lane >> 3, t = lane & 7;
    const u32 bars = (u32)__cvta_generic_to_shared(queue + QSLOTS * SLOT_BYTES);
    constexpr int U = 8;                                                                 // loads in flight per lane
    for (u32 k = 0;; k++) {
        const u32 s = k % QSLOTS, ph = (k / QSLOTS) & 1;
        mbar_wait(bars + s * 8, ph);
        const u8 *slot = queue + s * SLOT_BYTES;
        const uint4 *M = (const uint4 *)slot, *F = (const uint4 *)(slot + 512);
        const uint2 hdr = *(const uint2 *)(slot + 1024);
        if (hdr.x & Q_DONE) break;
        const u32 nm = hdr.x & 0xFF, nfr = (hdr.x >> 8) & 0xFF, total = hdr.y;
        for (u32 b0 = 0; b0 < nfr; b0 += 4) {
            const u32 idx = b0 + sub;
            const uint4 g = F[idx < nfr ? idx : 0];
            const u32 nw = idx < nfr ? g.y >> 24 : 0u;
            if (t < nw) make_ptr(g.x, g.y & 0xFFFFFFu)[t] = (u8)__byte_perm(g.z, g.w, t);
        }
        __syncwarp();
        for (u32 base = 0; base < total; base += 32 * U) {
            u8 *dst[U];
            u8 v[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const u32 b = base + u * 32 + lane;
                dst[u] = nullptr;
                if (b < total) {
                    u32 j = 0;
#pragma unroll
                    for (u32 step = 16; step; step >>= 1) {
                        const u32 c = j + step;
                        if (c < nm && M[c].w <= b) j = c;
                    }
                    const uint4 g = M[j];
                    const u32 i = b - g.w, d = g.z >> 16;
                    u8 *p = make_ptr(g.x, g.y);
                    u32 si = i;
                    if (si >= d) si %= d;
                    dst[u] = p + i;
                    v[u] = p[(i64)si - (i64)d];
                }
            }
#pragma unroll
            for (int u = 0; u < U; u++) if (dst[u]) *dst[u] = v[u];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + (QSLOTS + s) * 8);
    }
}

constexpr int NP = 4;                        // decoder/copier pairs per CTA; two CTAs per SM
// CTA layout: [rings: NP x 1 KiB][LUT_WORDS x 4][decoder tables: NP x WARP_BYTES][queues: NP x QUEUE_BYTES]
constexpr size_t SMEM_BYTES = (size_t)NP * RING_BYTES + LUT_WORDS * 4 + (size_t)NP * WARP_BYTES + (size_t)NP * QUEUE_BYTES;

__global__ void __launch_bounds__(NP * 64, 2)
inflate_fused_kernel(BatchArgs a) {
    extern __shared__ __align__(1024) u32 smem[];
    u32 *lut = smem + NP * RING_BYTES / 4;             // [0,32) length table, [32,64) distance table
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5, pair = warp % NP;
    u8 *queue = (u8 *)(lut + LUT_WORDS) + (size_t)NP * WARP_BYTES + (size_t)pair * QUEUE_BYTES;
    const u32 bars = (u32)__cvta_generic_to_shared(queue + QSLOTS * SLOT_BYTES);
    if (threadIdx.x < 32) { lut[threadIdx.x] = c_len_tab[threadIdx.x]; lut[32 + threadIdx.x] = c_dist_tab[threadIdx.x]; }
    if (warp < NP && lane == 0)
        for (int i = 0; i < 2 * QSLOTS; i++) mbar_init(bars + i * 8, 1);
    __syncthreads();
    if (warp >= NP) { copier_loop(queue, lane); return; }

    // ---------------------------------------------------------------------------------------------- decoder warp
    u8 *wbase = (u8 *)(lut + LUT_WORDS) + (size_t)warp * WARP_BYTES;
    u16 longsym[288];
    u8 longdst[32];
    LaneMem M;
    M.H = (u16 *)wbase + lane;
    M.W = (u32 *)(wbase + H_TOTAL * 32 * 2) + lane;
    M.longsym = longsym;
    M.longdst = longdst;
    u32 *ring = smem + warp * (RING_BYTES / 4) + lane;
    const u32 hlit = (u32)__cvta_generic_to_shared(M.H + H_LIT * 32);
    const u32 lt_mask = (1u << lane) - 1;
    u32 qk = 0;                                                                          // slots posted so far

    int status = SWC_OK, state = ST_DONE;
    bool have_unit = false, exhausted = false;
    u64 unit = 0;
    Span sp;
    Reader br;
    Emit em;
    BlockCtx bc;
    u64 cap64 = 0;
    u32 pend = 0;
    sp.origin = sp.ubeg = sp.uend = nullptr; sp.pos0 = sp.end = 0;
    br.lo = br.hi = 0; br.pos = 0; br.wend = 32; br.rptr = 0; br.wr = 0; br.nextc = 0; br.pre = make_uint4(0, 0, 0, 0);
    em.out = nullptr; em.op = 0; em.cap = 0; em.acc_lo = em.acc_hi = 0; em.nf = 0;
    for (;;) {
        if (state == ST_DONE) {
            if (have_unit) {                                                             // retire the finished unit
                if (status == SWC_OK && (u64)em.op > cap64) status = SWC_ERR_OUTPUT_OVERFLOW;
                a.consumed_bits[unit] = br.pos - sp.pos0;
                a.out_len[unit] = em.op;
                a.status[unit] = status;
                a.rec_count[unit] = 0;                                                   // nothing for lz_resolve_kernel to replay
                have_unit = false;
            }
            if (!exhausted) {
                unit = atomicAdd(a.ticket + 1, 1ull);
                if (unit >= a.n) {
                    exhausted = true;
                } else {
                    have_unit = true;
                    status = SWC_OK;
                    const u64 in_len = a.in_len[unit];
                    cap64 = a.out_cap[unit];
                    em.out = a.out_base + a.out_off[unit];
                    em.cap = cap64 > 0xFFFFFFF0ull ? 0xFFFFFFF0u : (u32)cap64;
                    em.op = 0;
                    const u32 bitskip = a.start_bits ? a.start_bits[unit] : 0;
                    sp.ubeg = a.in_base + a.in_off[unit];
                    sp.uend = sp.ubeg + in_len;
                    sp.origin = (const u8 *)((uintptr_t)sp.ubeg & ~(uintptr_t)15);
                    sp.pos0 = (u32)(sp.ubeg - sp.origin) * 8 + bitskip;
                    sp.end = sp.pos0 + (u32)(in_len * 8 - bitskip);
                    br.pos = sp.pos0;
                    if (in_len >= (1ull << 28)) status = SWC_ERR_UNSUPPORTED;            // bit positions are 32-bit here
                    else if (in_len * 8 - bitskip < 10) status = SWC_DEFLATE_WRONG_BLOCK_TYPE;      // Deflate.swift:36
                    else state = ST_HEADER;
                }
            }
        }
        if (!__any_sync(SWC_FULL, state != ST_DONE || have_unit || !exhausted)) break;
        if (state == ST_SYMBOLS) {
            br.topup(ring, sp);
            if (br.pos > sp.end) { status = SWC_DEFLATE_SYMBOL_NOT_FOUND; state = ST_DONE; }
        }
        // ---- fast: table lookups; a lane leaves the loop at its first non-literal ----
        if (state == ST_SYMBOLS) {
#pragma unroll
            for (int k = 0; k < KLIT; k++) {
                if (br.pos >= br.wend) br.advance();
                const u32 e = lds16(lut_addr(br.peek32() & ((1u << LB) - 1), hlit));
                if (e & E_NONLIT) { pend = e; state = ST_PARKED; break; }
                br.pos += e >> 8;
                em.literal(e);
            }
        }
        // ---- parked: lengths, distances, end of block, long codes ----
        Match mt; mt.len = 0; mt.dist = 0;
        if (state == ST_PARKED) {
            const int r = parked_step(br, em, M, bc, sp, lut, state, pend, mt);
            if (r) { status = r; state = ST_DONE; mt.len = 0; }
        }
        // ---- post this round's fragment and match to the copier warp ----
        const bool copy = mt.len != 0 && (u64)em.op + em.nf + mt.len <= em.cap;
        const u32 fm = __ballot_sync(SWC_FULL, em.nf != 0), cm = __ballot_sync(SWC_FULL, copy);
        if (fm | cm) {
            const u32 s = qk % QSLOTS, ph = (qk / QSLOTS) & 1;
            mbar_wait(bars + (QSLOTS + s) * 8, ph ^ 1);                                  // slot released by the copier (first lap: free)
            u8 *slot = queue + s * SLOT_BYTES;
            if (em.nf) {
                const u32 room = em.op < em.cap ? em.cap - em.op : 0u;
                const u32 nw = em.nf < room ? em.nf : room;                              // bytes that fit the capacity
                const u32 sh = 8 * (8 - em.nf);                                          // bring the first literal down to byte 0
                u32 lo, hi;
                if (sh >= 32) { lo = em.acc_hi >> (sh - 32); hi = 0; }
                else { lo = __funnelshift_r(em.acc_lo, em.acc_hi, sh); hi = em.acc_hi >> sh; }
                const u8 *dst = em.out + em.op;
                ((uint4 *)(slot + 512))[__popc(fm & lt_mask)] = make_uint4((u32)(uintptr_t)dst, (u32)((uintptr_t)dst >> 32) | (nw << 24), lo, hi);
                em.op += em.nf;
                em.nf = 0;
            }
            u32 incl = copy ? mt.len : 0u;                                                // prefix sums of the match lengths, lane order = entry order
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const u32 up = __shfl_up_sync(SWC_FULL, incl, d);
                if (lane >= (u32)d) incl += up;
            }
            const u32 total = __shfl_sync(SWC_FULL, incl, 31);
            if (copy) {
                const u8 *dst = em.out + em.op;
                ((uint4 *)slot)[__popc(cm & lt_mask)] = make_uint4((u32)(uintptr_t)dst, (u32)((uintptr_t)dst >> 32), mt.len | (mt.dist << 16), incl - mt.len);
                const u8 *src = dst - mt.dist;                                           // start the source sectors on their way into L2
                asm volatile("prefetch.global.L2 [%0];" ::"l"(src));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(src + (mt.len < mt.dist ? mt.len : mt.dist) - 1));
            }
            if (lane == 0) *(uint2 *)(slot + 1024) = make_uint2(__popc(cm) | (__popc(fm) << 8), total);
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + s * 8);
            qk++;
        }
        em.op += mt.len;
        // ---- header: lanes at a block boundary ----
        if (state == ST_HEADER) {
            HeaderBits hb;
            hb.sp = sp; hb.pos = br.pos;
            int next = ST_DONE;
            const int r = begin_block(hb, em, M, bc, next);
            br.pos = hb.pos;
            if (r) { status = r; state = ST_DONE; }
            else {
                state = next;
                if (next == ST_SYMBOLS) br.seek(ring, sp, hb.pos);
            }
        }
    }
    // tell the copier to stop
    {
        const u32 s = qk % QSLOTS, ph = (qk / QSLOTS) & 1;
        mbar_wait(bars + (QSLOTS + s) * 8, ph ^ 1);
        if (lane == 0) { *(uint2 *)(queue + s * SLOT_BYTES + 1024) = make_uint2(Q_DONE, 0); mbar_arrive(bars + s * 8); }
    }
}

}  // namespace k1f

int launch_fused(const BatchArgs &a, cudaStream_t stream) {
    int dev = 0;
    SWC_CUDA_TRY(cudaGetDevice(&dev));
    static bool configured[64] = {};
    static int num_sms[64] = {};
    if (!configured[dev & 63]) {
        SWC_CUDA_TRY(cudaFuncSetAttribute(k1f::inflate_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1f::SMEM_BYTES));
        SWC_CUDA_TRY(cudaDeviceGetAttribute(&num_sms[dev & 63], cudaDevAttrMultiProcessorCount, dev));
        configured[dev & 63] = true;
    }
    const u64 per_cta = k1f::NP * 32;                                        // decoder lanes per CTA
    u64 grid = (a.n + per_cta - 1) / per_cta;
    const u64 resident = (u64)num_sms[dev & 63] * 2;                         // persistent lanes: one CTA per resident slot
    if (grid > resident) grid = resident;
    k1f::inflate_fused_kernel<<<(unsigned)grid, k1f::NP * 64, k1f::SMEM_BYTES, stream>>>(a);
    count_launch();
    return SWC_OK;
}

}  // namespace inflate
}  // namespace swc
