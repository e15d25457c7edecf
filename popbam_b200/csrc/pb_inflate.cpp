// pb_inflate.cpp -- raw DEFLATE (RFC 1951) decoder for BGZF blocks, written for the host feeder.
//
// The reference inflates BGZF blocks with zlib's inflate() (bgzf.c:366-399), which is where 87 % of the
// feeder's time goes.  BGZF blocks are small (<= 64 KiB), whole in memory and
// with a known output size, which allows a simpler and faster decoder than a streaming one:
//   * 64-bit bit buffer refilled eight bytes at a time,
//   * one 11-bit primary table for literal/length codes (with sub-tables for the rare longer codes) and an
//     8-bit primary table for distance codes; every entry carries the symbol, its base value and the
//     number of extra bits, so a symbol is one lookup,
//   * two literals per lookup where both codes fit the primary index (read data is literal-heavy: few quality values,
//     short codes), so the serial chain shift -> index -> load -> shift is walked once for two bytes,
//   * a code's extra bits consumed together with the code (the entry holds the sum; the value is cut out of the bits as
//     they were before the shift), matches copied sixteen bytes at a time when they do not overlap closely.
// Output is byte-identical to zlib's (tests/test_host_feeder.py compares every block of the fixtures
// and randomised streams at all compression levels).
#include "pb_inflate.h"

#include <cstring>

namespace pbio {

namespace {

constexpr int kLitLenBits = 11;
constexpr int kDistBits = 8;
constexpr int kMaxCodeLen = 15;

// Table entry (32 bits):
//   bits  0..7   bits to consume: the code's length PLUS its extra bits (for a sub-table pointer: the primary bits)
//   bits  8..12  number of extra bits (length / distance symbols); literals: bit 8 = TWO literals (second one in bits 24..31)
//   bit   13     literal
//   bit   14     end of block
//   bit   15     sub-table pointer (value = offset of the sub-table, extra-bits field = its index width)
//   bits 16..31  literal byte / base length / base distance / sub-table offset
constexpr uint32_t kLiteral = 1u << 13, kEob = 1u << 14, kSub = 1u << 15, kDouble = 1u << 8;
// an index no code leads to: "end of block" and "literal" at once (and one bit long), which every path of the decoder rejects
// where it tests for the end of the block anyway -- no test of its own per symbol
constexpr uint32_t kInvalid = kEob | kLiteral | 1u;

const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

inline uint32_t reverse_bits(uint32_t v, int n) {
    uint32_t r = 0;
    for (int i = 0; i < n; ++i) { r = (r << 1) | (v & 1); v >>= 1; }
    return r;
}

// Builds a decode table from canonical code lengths.  `sym_entry(sym)` gives the entry without its length.
// Returns false for an over-subscribed or (non-trivially) incomplete code.
template <class F>
bool build_table(const uint8_t *lens, int n_syms, int primary_bits, uint32_t *table, int table_cap, F sym_entry) {
    int count[kMaxCodeLen + 1] = {0};
    for (int i = 0; i < n_syms; ++i) count[lens[i]]++;
    count[0] = 0;
    int used = 0;
    for (int l = 1; l <= kMaxCodeLen; ++l) used += count[l];
    const int primary_size = 1 << primary_bits;
    for (int i = 0; i < primary_size; ++i) table[i] = kInvalid;
    if (used == 0) return true;                       // no codes: only legal for an unused distance tree
    // canonical first codes
    uint32_t next_code[kMaxCodeLen + 2];
    uint32_t code = 0;
    int64_t left = 1;
    for (int l = 1; l <= kMaxCodeLen; ++l) {
        left <<= 1;
        left -= count[l];
        if (left < 0) return false;                   // over-subscribed
        code = (code + (uint32_t)count[l - 1]) << 1;
        next_code[l] = code;
    }
    if (left > 0 && !(used == 1)) return false;       // incomplete (a single code of length 1 is allowed)
    int sub_next = primary_size;
    // sub-table bookkeeping: for every primary prefix that needs one, its offset and width
    for (int sym = 0; sym < n_syms; ++sym) {
        const int l = lens[sym];
        if (!l) continue;
        const uint32_t c = next_code[l]++;
        const uint32_t rev = reverse_bits(c, l);
        const uint32_t e = sym_entry(sym);
        const uint32_t xb = (e & kLiteral) ? 0u : (e >> 8) & 31u;     // extra bits: consumed together with the code
        if (l <= primary_bits) {
            for (uint32_t i = rev; i < (uint32_t)primary_size; i += 1u << l) table[i] = e | ((uint32_t)l + xb);
        } else {
            const uint32_t prefix = rev & (uint32_t)(primary_size - 1);
            uint32_t p = table[prefix];
            if (!(p & kSub)) {
                // width of this sub-table: enough for the longest code sharing the prefix.  Codes are assigned
                // in increasing length, so scan the remaining lengths for the same prefix conservatively: use
                // the maximum code length present.
                int maxl = l;
                for (int l2 = kMaxCodeLen; l2 > l; --l2) if (count[l2]) { maxl = l2; break; }
                const int sub_bits = maxl - primary_bits;
                if (sub_next + (1 << sub_bits) > table_cap) return false;
                p = kSub | ((uint32_t)sub_bits << 8) | (uint32_t)primary_bits | ((uint32_t)sub_next << 16);
                table[prefix] = p;
                for (int i = 0; i < (1 << sub_bits); ++i) table[sub_next + i] = kInvalid;
                sub_next += 1 << sub_bits;
            }
            const int sub_bits = (int)((p >> 8) & 31);
            const uint32_t off = p >> 16;
            const uint32_t hi = rev >> primary_bits;             // the code's bits beyond the primary index
            for (uint32_t i = hi; i < (1u << sub_bits); i += 1u << (l - primary_bits)) table[off + i] = e | ((uint32_t)(l - primary_bits) + xb);
        }
    }
    return true;
}

struct Tables {
    uint32_t litlen[(1 << kLitLenBits) + 4608];    // + every possible sub-table (<= 286 prefixes x 16)
    uint32_t dist[(1 << kDistBits) + 3840];        // <= 30 prefixes x 128
};

bool build_litlen(const uint8_t *lens, int n, Tables &t) {
    if (!build_table(lens, n, kLitLenBits, t.litlen, (int)(sizeof t.litlen / 4), [](int sym) -> uint32_t {
        if (sym < 256) return kLiteral | ((uint32_t)sym << 16);
        if (sym == 256) return kEob;
        if (sym > 285) return kEob | kLiteral;        // invalid symbol: flagged, rejected when met
        return ((uint32_t)kLenBase[sym - 257] << 16) | ((uint32_t)kLenExtra[sym - 257] << 8);
    })) return false;
    // Two literals in one entry: where a literal's code leaves enough of the primary index for the whole code of another
    // literal, that one is determined by the index bits alone (prefix code) and the entry takes both.
    uint32_t one[1 << kLitLenBits];
    memcpy(one, t.litlen, sizeof one);
    for (uint32_t i = 0; i < (1u << kLitLenBits); ++i) {
        const uint32_t e1 = one[i];
        if ((e1 & (kLiteral | kEob | kSub)) != kLiteral) continue;
        const uint32_t l1 = e1 & 0xff;
        const uint32_t e2 = one[i >> l1];
        if ((e2 & (kLiteral | kEob | kSub)) != kLiteral) continue;
        const uint32_t l2 = e2 & 0xff;
        if (l1 + l2 > (uint32_t)kLitLenBits) continue;
        t.litlen[i] = kLiteral | kDouble | (l1 + l2) | (e1 & 0x00ff0000u) | ((e2 & 0x00ff0000u) << 8);
    }
    return true;
}
bool build_dist(const uint8_t *lens, int n, Tables &t) {
    return build_table(lens, n, kDistBits, t.dist, (int)(sizeof t.dist / 4), [](int sym) -> uint32_t {
        if (sym > 29) return kEob;                    // invalid distance symbol
        return ((uint32_t)kDistBase[sym] << 16) | ((uint32_t)kDistExtra[sym] << 8);
    });
}

struct BitReader {
    const uint8_t *p, *end;
    uint64_t buf = 0;
    int n = 0;          // valid bits in buf
    int over = 0;       // zero bytes appended past the end of the input (counted in n)
    // refill to at least 56 bits when the input allows; near the end byte by byte, then zeros
    inline void refill() {
        if (p + 8 <= end) {
            uint64_t w;
            memcpy(&w, p, 8);
            buf |= w << n;
            p += (63 - n) >> 3;                       // whole bytes that fit; n becomes 56 + (n & 7)
            n |= 56;
        } else {
            while (n <= 56) {
                if (p < end) buf |= (uint64_t)*p++ << n;
                else ++over;                          // past the end: zero bits (a valid stream never consumes them)
                n += 8;
            }
        }
    }
    inline uint32_t peek(int k) const { return (uint32_t)(buf & ((1ull << k) - 1)); }
    inline void drop(int k) { buf >>= k; n -= k; }
    inline uint32_t take(int k) { const uint32_t v = peek(k); drop(k); return v; }
};

const uint8_t kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

}  // namespace

bool inflate_raw(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len) {
    BitReader br{in, in + in_len};
    uint8_t *op = out, *const oend = out + out_len;
    static thread_local Tables tabs;
    static thread_local bool fixed_ready = false;
    static thread_local Tables fixed;
    bool final_block = false;
    while (!final_block) {
        br.refill();
        final_block = br.take(1) != 0;
        const uint32_t type = br.take(2);
        const Tables *T;
        if (type == 0) {                              // stored
            br.drop(br.n & 7);                        // to a byte boundary
            // give unread whole bytes back to the input pointer
            if ((br.n >> 3) < br.over) return false;  // the header itself ran past the end of the input
            const uint8_t *sp = br.p - ((br.n >> 3) - br.over);
            if (sp + 4 > br.end) return false;
            const uint32_t len = sp[0] | sp[1] << 8, nlen = sp[2] | sp[3] << 8;
            if ((len ^ 0xffffu) != nlen) return false;
            sp += 4;
            if (sp + len > br.end || op + len > oend) return false;
            memcpy(op, sp, len);
            op += len;
            br.p = sp + len; br.buf = 0; br.n = 0; br.over = 0;
            continue;
        } else if (type == 1) {                       // fixed Huffman codes
            if (!fixed_ready) {
                uint8_t l[288 + 32];
                for (int i = 0; i < 144; ++i) l[i] = 8;
                for (int i = 144; i < 256; ++i) l[i] = 9;
                for (int i = 256; i < 280; ++i) l[i] = 7;
                for (int i = 280; i < 288; ++i) l[i] = 8;
                for (int i = 0; i < 32; ++i) l[288 + i] = 5;
                if (!build_litlen(l, 288, fixed) || !build_dist(l + 288, 32, fixed)) return false;
                fixed_ready = true;
            }
            T = &fixed;
        } else if (type == 2) {                       // dynamic Huffman codes
            br.refill();
            const int hlit = (int)br.take(5) + 257, hdist = (int)br.take(5) + 1, hclen = (int)br.take(4) + 4;
            if (hlit > 286 || hdist > 30) return false;
            uint8_t cl[19] = {0};
            for (int i = 0; i < hclen; ++i) {
                if (br.n < 3) br.refill();
                cl[kClOrder[i]] = (uint8_t)br.take(3);
            }
            // the code-length code: 7-bit direct table
            uint32_t cltab[128 + 8];
            if (!build_table(cl, 19, 7, cltab, 136, [](int sym) -> uint32_t { return (uint32_t)sym << 16; })) return false;
            uint8_t lens[286 + 30 + 138];
            int i = 0;
            const int total = hlit + hdist;
            while (i < total) {
                br.refill();
                const uint32_t e = cltab[br.peek(7)];
                const int l = (int)(e & 0xff);
                if (e & kEob) return false;           // (no code leads here)
                br.drop(l);
                const int sym = (int)(e >> 16);
                if (sym < 16) lens[i++] = (uint8_t)sym;
                else {
                    int rep; uint8_t v = 0;
                    if (sym == 16) { if (!i) return false; v = lens[i - 1]; rep = 3 + (int)br.take(2); }
                    else if (sym == 17) rep = 3 + (int)br.take(3);
                    else rep = 11 + (int)br.take(7);
                    if (i + rep > total) return false;
                    memset(lens + i, v, (size_t)rep);
                    i += rep;
                }
            }
            if (!lens[256]) return false;             // no end-of-block code
            if (!build_litlen(lens, hlit, tabs) || !build_dist(lens + hlit, hdist, tabs)) return false;
            T = &tabs;
        } else return false;

        // ---- symbols
        const uint32_t *LL = T->litlen, *DD = T->dist;
        br.refill();                                  // >= 56 bits: enough for a length (15+5) and a distance (15+13) symbol
        uint32_t e = LL[br.peek(kLitLenBits)];        // the entry of the next symbol is always looked up one step ahead,
        for (;;) {                                    // so its load overlaps the copy of the match before it
            if (e & (kSub | kLiteral | kEob)) {       // (not a length: one test on the way to a match, the common symbol of read data)
                if (e & kSub) { br.drop(kLitLenBits); e = LL[(e >> 16) + br.peek((int)((e >> 8) & 31))]; }
                if (e & kLiteral) {
                    br.drop((int)(e & 0xff));             // (an index without a code: kInvalid, rejected here as a literal that is also the end of the block)
                    if (e & kEob) return false;           // invalid symbol 286/287
                    if ((size_t)(oend - op) >= 8) {
                        // one or two literals per entry (the second byte of a single one is overwritten by what follows);
                        // up to two more entries from the bits already in the buffer (3 x 11 + 15 <= 56)
                        uint16_t v = (uint16_t)(e >> 16);
                        memcpy(op, &v, 2); op += 1 + ((e >> 8) & 1u);
                        e = LL[br.peek(kLitLenBits)];
                        if ((e & (kLiteral | kEob | kSub)) == kLiteral) {
                            br.drop((int)(e & 0xff));
                            v = (uint16_t)(e >> 16); memcpy(op, &v, 2); op += 1 + ((e >> 8) & 1u);
                            e = LL[br.peek(kLitLenBits)];
                            if ((e & (kLiteral | kEob | kSub)) == kLiteral) {
                                br.drop((int)(e & 0xff));
                                v = (uint16_t)(e >> 16); memcpy(op, &v, 2); op += 1 + ((e >> 8) & 1u);
                            }
                        }
                    } else {
                        // the last bytes of the block, one at a time
                        const uint32_t cnt = 1 + ((e >> 8) & 1u);
                        if ((size_t)(oend - op) < cnt) return false;
                        *op++ = (uint8_t)(e >> 16);
                        if (cnt == 2) *op++ = (uint8_t)(e >> 24);
                    }
                    br.refill();
                    e = LL[br.peek(kLitLenBits)];
                    continue;
                }
                if (e & kEob) { br.drop((int)(e & 0xff)); break; }
            }
            const uint64_t saved = br.buf;            // (the extra bits of a length are taken from here: one shift consumes both)
            br.drop((int)(e & 0xff));
            const uint32_t xl = (e >> 8) & 31u;
            const uint32_t len = (e >> 16) + ((uint32_t)(saved >> ((e & 0xff) - xl)) & ((1u << xl) - 1u));
            uint32_t d = DD[br.peek(kDistBits)];
            if (d & (kSub | kEob)) {
                if (d & kSub) { br.drop(kDistBits); d = DD[(d >> 16) + br.peek((int)((d >> 8) & 31))]; }
                if (d & kEob) return false;           // invalid distance symbol, or an index without a code
            }
            // no refill needed here: the buffer held >= 56 bits when this symbol started (refilled before the look-ahead)
            // and a length (15 + 5) plus a distance (15 + 13) take at most 48
            const uint64_t saved_d = br.buf;
            br.drop((int)(d & 0xff));
            const uint32_t xd = (d >> 8) & 31u;
            const uint32_t dist = (d >> 16) + ((uint32_t)(saved_d >> ((d & 0xff) - xd)) & ((1u << xd) - 1u));
            if (dist > (size_t)(op - out) || len > (size_t)(oend - op)) return false;
            br.refill();
            e = LL[br.peek(kLitLenBits)];
            const uint8_t *src = op - dist;
            if (dist >= 16 && (size_t)(oend - op) >= len + 16) {
                // sixteen bytes at a time (the moves may run up to 15 bytes past the match: overwritten by what follows)
                struct W16 { uint64_t a, b; } w;
                memcpy(&w, src, 16); memcpy(op, &w, 16);
                for (uint32_t k = 16; k < len; k += 16) { memcpy(&w, src + k, 16); memcpy(op + k, &w, 16); }
                op += len;
            } else if (dist >= 8 && (size_t)(oend - op) >= len + 8) {
                uint64_t w;
                memcpy(&w, src, 8); memcpy(op, &w, 8);
                for (uint32_t k = 8; k < len; k += 8) { memcpy(&w, src + k, 8); memcpy(op + k, &w, 8); }
                op += len;
            } else {
                for (uint32_t k = 0; k < len; ++k) op[k] = src[k];
                op += len;
            }
        }
    }
    return op == oend;
}

}  // namespace pbio
